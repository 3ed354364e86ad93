"""persia-speedy wire format: `PersiaBatch.to_bytes()` / `get_embedding_from_bytes`.

The reference serialises `PersiaBatchImpl` (rust/persia-core/src/data.rs:34-41) with persia-speedy's derive
(`write_to_vec`, data.rs:256-258) and reads it back in `get_embedding_from_bytes` (lib.rs:400-407).  The encoding
rules below are the codec's own, each pinned by a vector of its test suite
(rust/persia-speedy/tests/serialization_tests.rs:600-1360; copied to tests/golden/speedy_vectors.json):

    little endian; bool / u8 one byte; usize as u64; f16 as its u16 bits
    Vec<T>, String, HashMap<K, V>: u32 length, then the elements (pairs)          (vec_u64, string, hashmap)
    Option<T>: u8 tag 0 / 1, then the value                                       (option_u16_some / _none)
    struct / tuple: the fields in declaration order, nothing in between           (derived_struct, tuple_u16_u16)
    enum: u32 tag = the variant's index (or explicit discriminant), then fields   (derived_enum_*, derived_simple_enum_*)
    SystemTime: u64 seconds + u32 nanoseconds since the epoch                     (system_time)

Structures (field order as declared in the reference):
    PersiaBatchImpl { non_id_type_features: Vec<TensorImpl>, id_type_features: EmbeddingTensor,
                      labels: Vec<TensorImpl>, meta_data: Option<Vec<u8>>, batch_id: Option<usize> }
    TensorImpl { storage: Storage, shape: Vec<usize>, stride: Vec<i64>, name: Option<String>, device: Device }
                                                                                  (persia-core/src/tensor.rs:229-236)
    Storage::CPU(CPUStorage) = tag 0; CPUStorage::{BOOL, F16, F32, F64, I8, I16, I32, I64, U8, U16, U32, U64}(Vec<_>)
                                                                                  (tensor.rs:97-111, 153-159)
    Device { device_type: DeviceType (CPU = 0, GPU = 1), device_id: Option<i32> }  (tensor.rs:170-180)
    EmbeddingTensor::{Null = 0, IDTypeFeature(IDTypeFeatureBatch) = 1, IDTypeFeatureRemoteRef(..) = 2}   (data.rs:15-20)
    IDTypeFeatureBatch { requires_grad: bool, batches: Vec<FeatureBatch>, enter_forward_id_buffer_time: Option<SystemTime>,
                         enter_post_forward_buffer_time: Option<SystemTime>, batcher_idx: Option<usize> }
                                                                                  (persia-common/src/lib.rs:117-126)
    FeatureBatch { feature_name: String, index_batch: Vec<SingleSignInFeatureBatch>, sample_num_signs: Vec<u32>,
                   hashed2index_batch_idx: HashMap<u64, i64>, batch_size: u16 }   (persia-common/src/lib.rs:29-43)
    SingleSignInFeatureBatch { sign: u64, in_which_batch_samples: Vec<(u16, u16)> } (persia-common/src/lib.rs:22-27)
    IDTypeFeatureRemoteRef { embedding_worker_addr: String, ref_id: u64, batcher_idx: usize }  (lib.rs:140-145)

The order of `index_batch` / `hashed2index_batch_idx` is hashbrown's iteration order in the reference (random per
process); this writer emits first-occurrence order.  Any order decodes to the same batch.
"""
import struct

import numpy as np

# CPUStorage variant index <-> numpy dtype (tensor.rs:97-111)
_STORAGE = [np.dtype(np.bool_), np.dtype(np.float16), np.dtype(np.float32), np.dtype(np.float64), np.dtype(np.int8),
            np.dtype(np.int16), np.dtype(np.int32), np.dtype(np.int64), np.dtype(np.uint8), np.dtype(np.uint16),
            np.dtype(np.uint32), np.dtype(np.uint64)]
_STORAGE_TAG = {dt: i for i, dt in enumerate(_STORAGE)}


class SpeedyError(RuntimeError):
    pass


class Writer:
    def __init__(self):
        self.parts = []

    def u8(self, v):
        self.parts.append(struct.pack("<B", v))

    def u16(self, v):
        self.parts.append(struct.pack("<H", v))

    def u32(self, v):
        self.parts.append(struct.pack("<I", v))

    def u64(self, v):
        self.parts.append(struct.pack("<Q", v))

    def i32(self, v):
        self.parts.append(struct.pack("<i", v))

    def i64(self, v):
        self.parts.append(struct.pack("<q", v))

    def raw(self, b):
        self.parts.append(bytes(b))

    def string(self, s):
        b = s.encode("utf-8")
        self.u32(len(b))
        self.raw(b)

    def array(self, a, dtype):
        """Vec<T> of a primitive: u32 length + little-endian elements."""
        a = np.ascontiguousarray(a, dtype=np.dtype(dtype).newbyteorder("<"))
        self.u32(a.size)
        self.raw(a.tobytes())

    def option(self, v, write):
        if v is None:
            self.u8(0)
        else:
            self.u8(1)
            write(v)

    def bytes(self):
        return b"".join(self.parts)


class Reader:
    def __init__(self, data):
        self.b = memoryview(bytes(data))
        self.at = 0

    def _take(self, n):
        if self.at + n > len(self.b):
            raise SpeedyError("unexpected end of input")  # speedy: Error::EndOfInput
        v = self.b[self.at:self.at + n]
        self.at += n
        return v

    def u8(self):
        return self._take(1)[0]

    def u16(self):
        return struct.unpack("<H", self._take(2))[0]

    def u32(self):
        return struct.unpack("<I", self._take(4))[0]

    def u64(self):
        return struct.unpack("<Q", self._take(8))[0]

    def i32(self):
        return struct.unpack("<i", self._take(4))[0]

    def i64(self):
        return struct.unpack("<q", self._take(8))[0]

    def string(self):
        return bytes(self._take(self.u32())).decode("utf-8")

    def array(self, dtype):
        dt = np.dtype(dtype).newbyteorder("<")
        n = self.u32()
        return np.frombuffer(self._take(n * dt.itemsize), dtype=dt).astype(np.dtype(dtype), copy=True)

    def option(self, read):
        tag = self.u8()
        if tag == 0:
            return None
        if tag != 1:
            raise SpeedyError("invalid Option tag")
        return read()

    def done(self):
        return self.at == len(self.b)


# ---- TensorImpl -----------------------------------------------------------------------------------------------------
def _write_tensor(w, name, arr):
    arr = np.asarray(arr)
    dt = np.dtype(arr.dtype.type)
    if dt not in _STORAGE_TAG:
        raise SpeedyError(f"Unsupport datatype of ndarray: {arr.dtype}")
    w.u32(0)                       # Storage::CPU
    w.u32(_STORAGE_TAG[dt])        # CPUStorage::<T>
    if dt == np.dtype(np.bool_):
        w.array(arr.reshape(-1).astype(np.uint8), np.uint8)
    else:
        w.array(arr.reshape(-1), dt)
    w.array(arr.shape, np.uint64)  # shape: Vec<usize>
    stride, acc = [], 1            # get_stride_by_shape (tensor.rs:215-227): row-major element strides
    for d in reversed(arr.shape):
        stride.append(acc)
        acc *= d
    w.array(list(reversed(stride)), np.int64)
    w.option(name, w.string)
    w.u32(0)                       # Device { device_type: CPU,
    w.u8(0)                        #          device_id: None }


def _read_tensor(r):
    if r.u32() != 0:
        raise SpeedyError("only CPU storage travels in a PersiaBatch")
    tag = r.u32()
    if tag >= len(_STORAGE):
        raise SpeedyError("invalid CPUStorage variant")
    dt = _STORAGE[tag]
    flat = r.array(np.uint8).astype(np.bool_) if dt == np.dtype(np.bool_) else r.array(dt)
    shape = tuple(int(x) for x in r.array(np.uint64))
    r.array(np.int64)  # stride (recomputed from the shape on this side)
    name = r.option(r.string)
    r.u32()
    r.option(r.i32)
    return name, flat.reshape(shape)


# ---- FeatureBatch -----------------------------------------------------------------------------------------------------
def _write_feature(w, name, feature):
    """feature: a uint64 ndarray (one id per sample) or a list of uint64 ndarrays (LIL).  FeatureBatch::new
    (persia-common/src/lib.rs:45-82) restated with first-occurrence order for the distinct signs."""
    rows = [np.array([v], np.uint64) for v in feature] if isinstance(feature, np.ndarray) else list(feature)
    where, signs, lists = {}, [], []
    for b, ids in enumerate(rows):
        for c, v in enumerate(np.asarray(ids, dtype=np.uint64).tolist()):
            k = where.get(v)
            if k is None:
                k = where[v] = len(signs)
                signs.append(v)
                lists.append([])
            lists[k].append((b, c))
    w.string(name)
    w.u32(len(signs))
    for s, occ in zip(signs, lists):
        w.u64(s)
        w.u32(len(occ))  # Vec<(u16, u16)>: the length counts pairs
        w.raw(np.asarray(occ, dtype="<u2").tobytes())
    w.array([len(r) for r in rows], np.uint32)
    w.u32(len(signs))
    for k, s in enumerate(signs):
        w.u64(s)
        w.i64(k)
    w.u16(len(rows))


def _read_feature(r):
    name = r.string()
    n = r.u32()
    batch_cells = {}
    for _ in range(n):
        sign = r.u64()
        pairs = r.u32()
        occ = np.frombuffer(r._take(pairs * 4), dtype="<u2").reshape(pairs, 2)
        for b, c in occ.tolist():
            batch_cells[(b, c)] = sign
    sample_num = r.array(np.uint32)
    m = r.u32()
    r._take(m * 16)  # hashed2index_batch_idx: derivable from index_batch
    bsz = r.u16()
    if len(sample_num) != bsz:
        raise SpeedyError("sample_num_signs does not match batch_size")
    rows = [np.array([batch_cells[(b, c)] for c in range(int(sample_num[b]))], np.uint64) for b in range(bsz)]
    if bsz and all(len(x) == 1 for x in rows):
        return name, np.concatenate(rows)
    return name, rows


# ---- PersiaBatchImpl --------------------------------------------------------------------------------------------------
def encode_batch(non_id, id_features, labels, meta, batch_id):
    """non_id / labels: [(name, ndarray)]; id_features: None | ("ids", requires_grad, [(name, feature)]) |
    ("ref", ref_id[, addr, batcher_idx])."""
    w = Writer()
    w.u32(len(non_id))
    for name, a in non_id:
        _write_tensor(w, name, a)
    if id_features is None:
        w.u32(0)
    elif id_features[0] == "ids":
        w.u32(1)
        w.u8(1 if id_features[1] else 0)
        w.u32(len(id_features[2]))
        for name, f in id_features[2]:
            _write_feature(w, name, f)
        w.u8(0)  # enter_forward_id_buffer_time: None
        w.u8(0)  # enter_post_forward_buffer_time: None
        w.u8(0)  # batcher_idx: None
    else:
        w.u32(2)
        w.string(id_features[2] if len(id_features) > 2 else "local")
        w.u64(int(id_features[1]))
        w.u64(int(id_features[3]) if len(id_features) > 3 else 0)
    w.u32(len(labels))
    for name, a in labels:
        _write_tensor(w, name, a)
    w.option(meta, lambda m: (w.u32(len(m)), w.raw(m)))
    w.option(batch_id, w.u64)
    return w.bytes()


def decode_batch(data):
    r = Reader(data)
    non_id = [_read_tensor(r) for _ in range(r.u32())]
    tag = r.u32()
    if tag == 0:
        idf = None
    elif tag == 1:
        requires_grad = bool(r.u8())
        feats = [_read_feature(r) for _ in range(r.u32())]
        for _ in range(2):
            r.option(lambda: (r.u64(), r.u32()))  # Option<SystemTime>
        r.option(r.u64)
        idf = ("ids", requires_grad, feats)
    elif tag == 2:
        addr = r.string()
        ref_id = r.u64()
        idx = r.u64()
        idf = ("ref", ref_id, addr, idx)
    else:
        raise SpeedyError("invalid EmbeddingTensor variant")
    labels = [_read_tensor(r) for _ in range(r.u32())]
    meta = r.option(lambda: bytes(r._take(r.u32())))
    batch_id = r.option(r.u64)
    if not r.done():
        raise SpeedyError("trailing bytes after PersiaBatchImpl")
    return non_id, idf, labels, meta, batch_id

"""CPU: the persia-speedy wire format (persia_b200/speedy.py) and the checkpoint codec (persia_b200/checkpoint.py)
against the codec's OWN golden vectors (rust/persia-speedy/tests/serialization_tests.rs, extracted to
tests/golden/speedy_vectors.json by tests/golden/make_speedy_vectors.py), plus `PersiaBatch.to_bytes()` /
`get_embedding_from_bytes` decoding of a PersiaBatchImpl laid out by hand in the reference's field order."""
import json
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from persia_b200 import speedy as SP  # noqa: E402

V = json.load(open(os.path.join(ROOT, "tests", "golden", "speedy_vectors.json")))["vectors"]


def _le(name):
    return bytes(V[name]["le"])


def test_primitive_rules_match_the_codecs_vectors():
    w = SP.Writer()
    w.array([10, 11], np.uint8)
    assert w.bytes() == _le("vec_u8")                      # Vec<u8>: u32 length prefix
    for name, dt in (("vec_u16", np.uint16), ("vec_u32", np.uint32), ("vec_u64", np.uint64)):
        w = SP.Writer()
        w.array([10, 11], dt)
        assert w.bytes() == _le(name)
        assert SP.Reader(_le(name)).array(dt).tolist() == [10, 11]
    for name, fn, val in (("u16", "u16", 33), ("u32", "u32", 33), ("u64", "u64", 33), ("usize", "u64", 33),
                          ("i32", "i32", -33), ("i64", "i64", -33), ("bool_true", "u8", 1), ("bool_false", "u8", 0)):
        w = SP.Writer()
        getattr(w, fn)(val)
        assert w.bytes() == _le(name), name
        assert getattr(SP.Reader(_le(name)), fn)() == val
    assert struct.pack("<f", 8388610.0) == _le("f32") and struct.pack("<d", 8388610.0) == _le("f64")
    w = SP.Writer()
    w.string("Hello")
    assert w.bytes() == _le("string") and SP.Reader(_le("string")).string() == "Hello"
    w = SP.Writer()
    w.u16(10)
    w.u16(11)
    assert w.bytes() == _le("tuple_u16_u16")               # tuples / structs: fields back to back
    w = SP.Writer()
    w.option(10, w.u16)
    assert w.bytes() == _le("option_u16_some")
    w = SP.Writer()
    w.option(None, w.u16)
    assert w.bytes() == _le("option_u16_none")
    r = SP.Reader(_le("option_u16_some"))
    assert r.option(r.u16) == 10
    w = SP.Writer()                                        # HashMap<u16, bool>: u32 length, then (key, value) pairs
    w.u32(1)
    w.u16(10)
    w.u8(1)
    assert w.bytes() == _le("hashmap")
    assert _le("system_time") == struct.pack("<QI", 0, 0)  # SystemTime: u64 seconds + u32 nanoseconds
    # enums: u32 tag = variant index, or the explicit discriminant (B = 10, C = 11 there)
    assert _le("derived_simple_enum_a") == struct.pack("<I", 0) and _le("derived_simple_enum_b") == struct.pack("<I", 10)
    assert _le("derived_simple_enum_c") == struct.pack("<I", 11)
    assert _le("derived_enum_unit_variant") == struct.pack("<I", 0)
    assert _le("derived_enum_tuple_variant") == struct.pack("<IBHI", 1, 10, 20, 30)
    assert _le("derived_enum_struct_variant") == struct.pack("<IBHI", 2, 100, 200, 300)
    assert _le("derived_struct") == struct.pack("<BHI", 1, 2, 3)


def test_checkpoint_codec_uses_the_same_rules():
    """checkpoint.py's `.emb` layout is built from exactly these primitives: usize as u64, u32 indices, u8 flags, u32 length
    prefixes, f32 little endian (persia-embedding-holder array_linked_list.rs:137-213, emb_entry.rs:17-25)."""
    from persia_b200 import checkpoint as CK

    signs = np.array([33, 34], np.uint64)
    ent = np.array([[8388610.0, 1.0], [2.0, 3.0]], np.float32)
    blob = CK.encode_list(signs, ent, 2)
    r = SP.Reader(blob)
    assert r.u64() == 2                                                   # count: usize  (vector `usize`)
    assert (r.u32(), r.u32(), r.u32(), r.u32()) == (1, 2, 0, 0)           # first, last, free, end: u32
    assert r.u32() == 2                                                   # Vec<node>: u32 length (vector `vec_u32`)
    node0 = (r.u32(), r.u32(), r.u8())                                    # next, prev, Option tag
    assert node0 == (2, 0, 1)
    inner = r.array(np.float32)                                           # Vec<f32>
    assert inner.tobytes()[:4] == _le("f32") and inner.tolist() == [8388610.0, 1.0]
    assert (r.u64(), r.u64()) == (2, 33)                                  # embedding_dim: usize, sign: u64
    got_signs, got_dims, got_ent = CK.decode_list(blob)
    assert got_signs.tolist() == [33, 34] and np.array_equal(np.stack(got_ent), ent) and list(got_dims) == [2, 2]


def _hand_laid_batch():
    """A PersiaBatchImpl written field by field with struct.pack only (independent of speedy.py), index_batch in an
    order a hashbrown map could produce (not first-occurrence order)."""
    def tensor(tag, payload, n, shape, name):
        b = struct.pack("<II", 0, tag) + struct.pack("<I", n) + payload
        b += struct.pack("<I", len(shape)) + b"".join(struct.pack("<Q", s) for s in shape)
        stride = [1] * len(shape)
        for i in range(1, len(shape)):
            stride[len(shape) - i - 1] = stride[len(shape) - i] * shape[len(shape) - i]
        b += struct.pack("<I", len(shape)) + b"".join(struct.pack("<q", s) for s in stride)
        b += b"\x01" + struct.pack("<I", len(name)) + name.encode() if name else b"\x00"
        return b + struct.pack("<I", 0) + b"\x00"

    dense = np.arange(6, dtype=np.float32).reshape(3, 2)
    label = np.array([[1], [0], [1]], np.int64)
    out = struct.pack("<I", 1) + tensor(2, dense.tobytes(), 6, (3, 2), "dense")
    out += struct.pack("<I", 1) + b"\x01" + struct.pack("<I", 2)          # IDTypeFeature, requires_grad, 2 features
    # feature "a": one id per sample: ids 7, 9, 7  -> signs listed as 9 then 7
    out += struct.pack("<I", 1) + b"a" + struct.pack("<I", 2)
    out += struct.pack("<Q", 9) + struct.pack("<I", 1) + struct.pack("<HH", 1, 0)
    out += struct.pack("<Q", 7) + struct.pack("<I", 2) + struct.pack("<HHHH", 0, 0, 2, 0)
    out += struct.pack("<I", 3) + struct.pack("<III", 1, 1, 1)
    out += struct.pack("<I", 2) + struct.pack("<Qq", 9, 0) + struct.pack("<Qq", 7, 1) + struct.pack("<H", 3)
    # feature "b": LIL [[5, 6], [], [6]]
    out += struct.pack("<I", 1) + b"b" + struct.pack("<I", 2)
    out += struct.pack("<Q", 6) + struct.pack("<I", 2) + struct.pack("<HHHH", 0, 1, 2, 0)
    out += struct.pack("<Q", 5) + struct.pack("<I", 1) + struct.pack("<HH", 0, 0)
    out += struct.pack("<I", 3) + struct.pack("<III", 2, 0, 1)
    out += struct.pack("<I", 2) + struct.pack("<Qq", 6, 0) + struct.pack("<Qq", 5, 1) + struct.pack("<H", 3)
    out += b"\x00\x00\x00"                                                 # two Option<SystemTime> and batcher_idx: None
    out += struct.pack("<I", 1) + tensor(7, label.tobytes(), 3, (3, 1), None)
    out += b"\x01" + struct.pack("<I", 2) + b"hi"                         # meta_data: Some(b"hi")
    out += b"\x01" + struct.pack("<Q", 42)                                 # batch_id: Some(42)
    return out, dense, label


def test_decode_reference_layout_and_round_trip():
    blob, dense, label = _hand_laid_batch()
    non_id, idf, labels, meta, batch_id = SP.decode_batch(blob)
    assert non_id[0][0] == "dense" and np.array_equal(non_id[0][1], dense)
    assert labels[0][0] is None and np.array_equal(labels[0][1], label) and labels[0][1].dtype == np.int64
    assert meta == b"hi" and batch_id == 42
    kind, requires_grad, feats = idf
    assert kind == "ids" and requires_grad is True
    assert feats[0][0] == "a" and feats[0][1].tolist() == [7, 9, 7]
    assert feats[1][0] == "b" and [x.tolist() for x in feats[1][1]] == [[5, 6], [], [6]]
    # our writer's bytes decode to the same batch (the order of the distinct signs is free)
    again = SP.decode_batch(SP.encode_batch(non_id, idf, labels, meta, batch_id))
    assert again[1][2][0][1].tolist() == [7, 9, 7] and [x.tolist() for x in again[1][2][1][1]] == [[5, 6], [], [6]]
    assert np.array_equal(again[0][0][1], dense) and again[3] == b"hi" and again[4] == 42
    with pytest.raises(SP.SpeedyError):
        SP.decode_batch(blob[:-3])
    with pytest.raises(SP.SpeedyError):
        SP.decode_batch(blob + b"\x00")


def test_persia_batch_to_bytes_is_speedy():
    from persia_b200 import persia_core as PC

    b = PC.PersiaBatch()
    b.add_non_id_type_feature(np.ones((2, 3), np.float32), np.dtype(np.float32), "d")
    b.add_id_type_feature_with_single_id(np.array([3, 3], np.uint64), "slot")
    b.add_label(np.zeros((2, 1), np.float32), np.dtype(np.float32), "y")
    b.add_meta(b"m")
    b.converted_id_type_features2embedding_tensor(True)
    blob = b.to_bytes()
    assert blob[:4] == struct.pack("<I", 1) and blob[4:12] == struct.pack("<II", 0, 2)  # 1 tensor, Storage::CPU, CPUStorage::F32
    p = PC.PersiaBatch._from_bytes(blob)
    assert p.embedding_tensor[0] == "ids" and p.embedding_tensor[1] is True
    assert p.embedding_tensor[2][0][0] == "slot" and p.embedding_tensor[2][0][1].tolist() == [3, 3]
    assert p.meta_data == b"m" and p.non_id_type_features[0][0] == "d"
    with pytest.raises(RuntimeError):
        PC.PersiaBatch._from_bytes(b"\x01\x02")

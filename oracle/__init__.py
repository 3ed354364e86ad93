"""ctypes front-end of the CPU oracle (oracle/persia_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  persia_b200/ never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpersia_oracle.so")

SGD, ADAGRAD, ADAGRAD_VW, ADAM = 0, 1, 2, 3


def build(force=False):
    src = os.path.join(_HERE, "persia_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpersia_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u64, u32, f32, i32, vp = C.c_uint64, C.c_uint32, C.c_float, C.c_int, C.c_void_p
    sig = {
        "po_farmhash64": (u64, [u64]),
        "po_farmhash64_many": (None, [vp, u64, vp]),
        "po_shard_of": (None, [vp, u64, u64, vp]),
        "po_add_prefix": (None, [vp, u64, u32, u64]),
        "po_index_prefix": (u64, [u32, u32]),
        "po_set_rsqrt_exact": (None, [i32]),
        "po_init_row": (None, [u64, u32, f32, f32, vp]),
        "po_f32_to_f16": (None, [vp, u64, vp]),
        "po_add_assign": (None, [vp, vp, u64]),
        "po_weight_bound": (None, [vp, u64, f32]),
        "po_optim_require_space": (u32, [i32, u32]),
        "po_optim_state_init": (None, [i32, f32, vp, u32]),
        "po_optim_update": (None, [i32, f32, f32, f32, f32, f32, f32, f32, f32, vp, u32, vp, u32]),
        "po_evmap_new": (vp, [u64]),
        "po_evmap_free": (None, [vp]),
        "po_evmap_insert": (None, [vp, u64]),
        "po_evmap_get_refresh": (i32, [vp, u64]),
        "po_evmap_len": (u64, [vp]),
        "po_fb_new": (vp, [vp, vp, u32]),
        "po_fb_free": (None, [vp]),
        "po_fb_hashstack": (None, [vp, u32, u64]),
        "po_fb_add_prefix": (None, [vp, u32, u64]),
        "po_fb_num_unique": (u32, [vp]),
        "po_fb_num_occ": (u32, [vp]),
        "po_fb_export": (None, [vp, vp, vp, vp, vp, vp]),
        "po_worker_new": (vp, [u32, u32, u64, u32, u32]),
        "po_worker_free": (None, [vp]),
        "po_worker_set_slot": (None, [vp, u32, u32, i32, i32, u32, u32, u64, u64]),
        "po_worker_configure": (None, [vp, f32, f32, f32, i32, f32]),
        "po_worker_set_optimizer": (None, [vp, i32, f32, f32, f32, f32, f32, f32, f32]),
        "po_worker_set_faithful_miss": (None, [vp, i32]),
        "po_worker_set_embedding": (None, [vp, vp, u64, vp, u32, u32]),
        "po_worker_get_entry": (i32, [vp, u64, vp, u32]),
        "po_worker_ps_len": (u64, [vp, u32]),
        "po_worker_grad_miss": (u64, [vp]),
        "po_ps_lookup": (i32, [vp, u32, vp, vp, u64, i32, vp]),
        "po_ps_update": (i32, [vp, u32, vp, vp, u64, vp, u64]),
        "po_ctx_new": (vp, []),
        "po_ctx_free": (None, [vp]),
        "po_ctx_num_unique": (u32, [vp, u32]),
        "po_ctx_signs": (None, [vp, u32, vp]),
        "po_worker_forward": (i32, [vp, vp, vp, u32, i32, vp, vp]),
        "po_worker_backward": (i32, [vp, vp, vp, i32, vp, vp, vp]),
        "po_worker_forward_raw": (i32, [vp, u32, vp, vp, u32, i32, vp, vp, vp, vp, vp]),
        "po_worker_backward_raw": (i32, [vp, u32, vp, vp, i32, f32, i32, vp]),
        "po_worker_bench": (C.c_double, [vp, vp, vp, u32, u64, u32, vp, u32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ---- leaf functions ---------------------------------------------------------
def farmhash64(x):
    x = _c(x, np.uint64)
    out = np.empty_like(x)
    lib().po_farmhash64_many(_p(x), x.size, _p(out))
    return out


def shard_of(signs, R):
    signs = _c(signs, np.uint64)
    out = np.empty(signs.size, np.uint32)
    lib().po_shard_of(_p(signs), signs.size, R, _p(out))
    return out


def add_prefix(ids, prefix_bit, prefix):
    s = _c(ids, np.uint64).copy()
    lib().po_add_prefix(_p(s), s.size, prefix_bit, prefix)
    return s


def index_prefix(group_index, prefix_bit=8):
    return int(lib().po_index_prefix(group_index, prefix_bit))


def set_rsqrt_exact(on):
    lib().po_set_rsqrt_exact(int(bool(on)))


def init_row(sign, dim, lo, hi):
    out = np.empty(dim, np.float32)
    lib().po_init_row(int(sign), dim, lo, hi, _p(out))
    return out


def f32_to_f16(a):
    a = _c(a, np.float32)
    out = np.empty(a.shape, np.uint16)
    lib().po_f32_to_f16(_p(a), a.size, _p(out))
    return out.view(np.float16)


class Optim:
    """persia-common/src/optim.rs Optimizable on a bare entry."""

    def __init__(self, kind, lr=0.01, wd=0.0, mom=1.0, init_acc=0.01, eps=1e-10, b1=0.9, b2=0.999):
        self.kind, self.lr, self.wd, self.mom, self.init_acc, self.eps, self.b1, self.b2 = (
            kind, lr, wd, mom, init_acc, eps, b1, b2)

    def require_space(self, dim):
        return int(lib().po_optim_require_space(self.kind, dim))

    def new_entry(self, emb):
        emb = _c(emb, np.float32)
        dim = emb.size
        e = np.zeros(dim + self.require_space(dim), np.float32)
        e[:dim] = emb
        lib().po_optim_state_init(self.kind, self.init_acc, _p(e), dim)
        return e

    def update(self, entry, grad, dim, b1p=0.0, b2p=0.0):
        grad = _c(grad, np.float32)
        lib().po_optim_update(self.kind, self.lr, self.wd, self.mom, self.eps, self.b1, self.b2, b1p, b2p,
                              _p(entry), entry.size, _p(grad), dim)


def weight_bound(emb, b):
    lib().po_weight_bound(_p(emb), emb.size, b)


class EvictionMap:
    def __init__(self, cap):
        self.h = lib().po_evmap_new(cap)

    def insert(self, k):
        lib().po_evmap_insert(self.h, k)

    def get_refresh(self, k):
        return bool(lib().po_evmap_get_refresh(self.h, k))

    def __len__(self):
        return int(lib().po_evmap_len(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().po_evmap_free(self.h)
            self.h = None


def lil_to_csr(batch):
    """List[List[int]] (one slot) -> (ids u64, row_off u32[B+1])."""
    off = np.zeros(len(batch) + 1, np.uint32)
    off[1:] = np.cumsum([len(x) for x in batch])
    ids = np.array([i for x in batch for i in x], dtype=np.uint64)
    return ids, off


class FeatureBatch:
    """persia-common/src/lib.rs:45-82, uniques in first-occurrence order."""

    def __init__(self, ids, row_off):
        self.ids, self.row_off = _c(ids, np.uint64), _c(row_off, np.uint32)
        self.h = lib().po_fb_new(_p(self.ids), _p(self.row_off), self.row_off.size - 1)
        if not self.h:
            raise RuntimeError("batch size cannot be larger than 65535")

    def hashstack(self, rounds, size):
        lib().po_fb_hashstack(self.h, rounds, size)

    def add_prefix(self, prefix_bit, prefix):
        lib().po_fb_add_prefix(self.h, prefix_bit, prefix)

    def export(self):
        U, n, B = lib().po_fb_num_unique(self.h), lib().po_fb_num_occ(self.h), self.row_off.size - 1
        signs, seg = np.empty(U, np.uint64), np.empty(U + 1, np.uint32)
        os_, oc, sns = np.empty(n, np.uint16), np.empty(n, np.uint16), np.empty(B, np.uint32)
        lib().po_fb_export(self.h, _p(signs), _p(seg), _p(os_), _p(oc), _p(sns))
        return signs, seg, os_, oc, sns

    def __del__(self):
        if getattr(self, "h", None):
            lib().po_fb_free(self.h)
            self.h = None


class SlotCfg:
    def __init__(self, dim, summation=True, sqrt_scaling=False, sample_fixed_size=10, hs_rounds=0, hs_size=0,
                 prefix=0):
        self.dim, self.summation, self.sqrt_scaling = dim, summation, sqrt_scaling
        self.sample_fixed_size, self.hs_rounds, self.hs_size, self.prefix = sample_fixed_size, hs_rounds, hs_size, prefix


class Worker:
    """The reference's embedding worker + R parameter servers, in process."""

    def __init__(self, slots, n_ps=1, capacity_per_ps=1 << 30, n_internal_shards=1, prefix_bit=8):
        self.slots, self.R = list(slots), n_ps
        L = lib()
        self.h = L.po_worker_new(len(slots), n_ps, capacity_per_ps, n_internal_shards, prefix_bit)
        for i, s in enumerate(self.slots):
            L.po_worker_set_slot(self.h, i, s.dim, int(s.summation), int(s.sqrt_scaling), s.sample_fixed_size,
                                 s.hs_rounds, s.hs_size, s.prefix)

    def configure(self, lo=-0.01, hi=0.01, admit_p=1.0, enable_wb=True, wb=10.0):
        lib().po_worker_configure(self.h, lo, hi, admit_p, int(enable_wb), wb)

    def set_optimizer(self, o):
        lib().po_worker_set_optimizer(self.h, o.kind, o.lr, o.wd, o.mom, o.init_acc, o.eps, o.b1, o.b2)

    def set_faithful_miss(self, on):
        lib().po_worker_set_faithful_miss(self.h, int(on))

    def set_embedding(self, signs, entries, dim):
        signs, entries = _c(signs, np.uint64), _c(entries, np.float32)
        lib().po_worker_set_embedding(self.h, _p(signs), signs.size, _p(entries), dim, entries.shape[1])

    def get_entry(self, sign, max_len=4096):
        out = np.empty(max_len, np.float32)
        n = lib().po_worker_get_entry(self.h, int(sign), _p(out), max_len)
        return out[:n].copy() if n > 0 else None

    def ps_len(self, r=0):
        return int(lib().po_worker_ps_len(self.h, r))

    def grad_miss(self):
        return int(lib().po_worker_grad_miss(self.h))

    def ps_lookup(self, r, signs, dims, training):
        signs, dims = _c(signs, np.uint64), _c(dims, np.uint32)
        out = np.empty(int(dims.sum()), np.float32)
        rc = lib().po_ps_lookup(self.h, r, _p(signs), _p(dims), signs.size, int(training), _p(out))
        if rc != 0:
            raise RuntimeError(f"ps lookup failed rc={rc}")
        return out

    def ps_update(self, r, signs, dims, grads):
        signs, dims, grads = _c(signs, np.uint64), _c(dims, np.uint32), _c(grads, np.float32)
        rc = lib().po_ps_update(self.h, r, _p(signs), _p(dims), signs.size, _p(grads), grads.size)
        if rc != 0:
            raise RuntimeError(f"ps update failed rc={rc}")

    def forward(self, ids, row_off, B, training=True, keep_ctx=True):
        """ids: flat u64 slot-major; row_off: u32[S*B+1].  Returns (list of [B,dim] f16 per slot, ctx)."""
        ids, row_off = _c(ids, np.uint64), _c(row_off, np.uint32)
        assert row_off.size == len(self.slots) * B + 1
        tot = sum(B * s.dim for s in self.slots)
        out = np.empty(tot, np.uint16)
        ctx = lib().po_ctx_new() if keep_ctx else None
        rc = lib().po_worker_forward(self.h, _p(ids), _p(row_off), B, int(training), _p(out), ctx)
        if rc != 0:
            raise RuntimeError(f"oracle forward failed rc={rc}")
        res, o = [], 0
        for s in self.slots:
            res.append(out[o:o + B * s.dim].view(np.float16).reshape(B, s.dim))
            o += B * s.dim
        return res, ctx

    def ctx_signs(self, ctx, slot):
        U = lib().po_ctx_num_unique(ctx, slot)
        out = np.empty(U, np.uint64)
        lib().po_ctx_signs(ctx, slot, _p(out))
        return out

    def backward(self, ctx, grads, scale=None, skip=None, free_ctx=True):
        """grads: list of [B,dim] arrays (all f16 or all f32).  Returns per-slot status list."""
        is_f16 = grads[0].dtype == np.float16
        flat = np.concatenate([_c(g, np.float16 if is_f16 else np.float32).reshape(-1) for g in grads])
        S = len(self.slots)
        sc = _c(scale, np.float32) if scale is not None else None
        sk = _c(skip, np.int32) if skip is not None else None
        status = np.zeros(S, np.int32)
        rc = lib().po_worker_backward(self.h, ctx, _p(flat), int(is_f16), _p(sc), _p(sk), _p(status))
        if free_ctx:
            lib().po_ctx_free(ctx)
        if rc != 0:
            raise RuntimeError(f"oracle backward failed rc={rc}")
        return status.tolist()

    def forward_raw(self, slot, ids, row_off, B, training=True, keep_ctx=True):
        """One raw slot.  ids: the slot's flat u64 ids; row_off: u32[B+1].
        Returns (table f16 [U+1, dim], index i64 [B*fixed], non_empty i64, sample_id_num u32 [B], ctx)."""
        ids, row_off = _c(ids, np.uint64), _c(row_off, np.uint32)
        s = self.slots[slot]
        table = np.zeros((ids.size + 1) * s.dim, np.uint16)
        index = np.zeros(B * s.sample_fixed_size, np.int64)
        num = np.zeros(max(B, 1), np.uint32)
        U = np.zeros(1, np.uint32)
        ctx = lib().po_ctx_new() if keep_ctx else None
        rc = lib().po_worker_forward_raw(self.h, slot, _p(ids), _p(row_off), B, int(training), _p(table), _p(index),
                                         _p(num), _p(U), ctx)
        if rc != 0:
            raise RuntimeError(f"oracle raw forward failed rc={rc}")
        U = int(U[0])
        non_empty = np.nonzero(index)[0].astype(np.int64)  # persia-core forward.rs:336-347
        return table[:(U + 1) * s.dim].view(np.float16).reshape(U + 1, s.dim), index, non_empty, num[:B], ctx

    def backward_raw(self, slot, ctx, grad, scale=1.0, skip=False, free_ctx=True):
        """grad: [U, dim] f32 or f16 (None with skip=True).  Returns the slot status (0 applied, 1 skipped, 2 NaN)."""
        is_f16 = grad is not None and grad.dtype == np.float16
        g = _c(grad, np.float16 if is_f16 else np.float32) if grad is not None else None
        status = np.zeros(1, np.int32)
        rc = lib().po_worker_backward_raw(self.h, slot, ctx, _p(g), int(is_f16), float(scale), int(skip or g is None),
                                          _p(status))
        if free_ctx:
            lib().po_ctx_free(ctx)
        if rc != 0:
            raise RuntimeError(f"oracle raw backward failed rc={rc}")
        return int(status[0])

    def bench(self, ids_batches, row_off, B, grads_f16, n_threads):
        """ids_batches: [n_batches, ids_per_batch] u64.  Returns wall seconds for fwd+bwd of all batches."""
        ids_batches, row_off = _c(ids_batches, np.uint64), _c(row_off, np.uint32)
        g = np.concatenate([_c(x, np.float16).reshape(-1) for x in grads_f16])
        return float(lib().po_worker_bench(self.h, _p(ids_batches), _p(row_off), B, ids_batches.shape[1],
                                           ids_batches.shape[0], _p(g), n_threads))

    def __del__(self):
        if getattr(self, "h", None):
            lib().po_worker_free(self.h)
            self.h = None

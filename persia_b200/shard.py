"""EmbeddingShard / BatchContext: thin torch-facing wrappers over the C ABI.

`EmbeddingShard` is one embedding-parameter-server replica resident in one GPU's HBM
(reference: rust/persia-embedding-server/src/embedding_parameter_service/mod.rs +
rust/persia-embedding-holder).  `BatchContext` is what the embedding worker keeps between the forward
and the backward of one batch (embedding_worker_service/mod.rs:1087-1119).  torch only owns device
memory and the current stream here; all compute is in libpersia_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import native as N


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _as_i64_bits(t):
    """uint64 ids travel as int64 bit patterns inside torch (torch has no full uint64 support)."""
    if t.dtype == torch.uint64:
        t = t.view(torch.int64)
    assert t.dtype == torch.int64, "ids must be uint64/int64 bit patterns"
    return t.contiguous()


class EmbeddingShard:
    def __init__(self, dim, capacity, device=0):
        self.lib = N.load()
        self.dim, self.capacity = int(dim), int(capacity)
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        h = C.c_void_p()
        cfg = N.TableCfg(self.dim, self.capacity)
        N.check(self.lib.pb_table_create(self.device.index or 0, C.byref(cfg), C.byref(h)))
        self.h = h
        self._entry_len = None

    # register_optimizer (PS mod.rs:429-438)
    def set_optimizer(self, kind, lr=0.01, wd=0.0, g_square_momentum=1.0, initialization=0.01, eps=1e-10,
                      beta1=0.9, beta2=0.999):
        cfg = N.OptimCfg(kind, lr, wd, g_square_momentum, initialization, eps, beta1, beta2)
        N.check(self.lib.pb_table_set_optimizer(self.h, C.byref(cfg)))
        self._entry_len = None

    # configure (PS mod.rs:440-451)
    def configure(self, init_lower=-0.01, init_upper=0.01, admit_probability=1.0, enable_weight_bound=True,
                  weight_bound=10.0):
        cfg = N.HyperCfg(init_lower, init_upper, admit_probability, int(enable_weight_bound), weight_bound)
        N.check(self.lib.pb_table_configure(self.h, C.byref(cfg)))

    def set_eviction(self, check_every=4, low_water=None, target_free=None, keep_batches=2):
        """Recency-based capacity policy (EvictionMap semantics, batch-granular); see persia_b200.h."""
        low = (self.capacity // 16 if low_water is None else low_water) if check_every else 0
        tgt = (self.capacity // 8 if target_free is None else target_free) if check_every else 0
        N.check(self.lib.pb_table_set_eviction(self.h, check_every, low, tgt, keep_batches))

    @property
    def entry_len(self):
        if self._entry_len is None:
            v = C.c_uint32()
            N.check(self.lib.pb_table_entry_len(self.h, C.byref(v)))
            self._entry_len = v.value
        return self._entry_len

    def counters(self):
        out = (C.c_uint64 * 5)()
        N.check(self.lib.pb_table_counters(self.h, C.byref(out), _stream(self.device)))
        return {"admitted": out[0], "lookup_miss": out[1], "gradient_id_miss": out[2], "capacity_refused": out[3],
                "wait_errors": out[4]}

    def __len__(self):
        return int(self.counters()["admitted"])

    def clear(self):
        N.check(self.lib.pb_table_clear(self.h, _stream(self.device)))

    # lookup_mixed (PS mod.rs:344-357)
    def lookup(self, signs, training=True, out=None):
        signs = _as_i64_bits(signs)
        n = signs.numel()
        if out is None:
            out = torch.empty((n, self.dim), dtype=torch.float32, device=self.device)
        N.check(self.lib.pb_lookup(self.h, _ptr(signs), n, int(training), _ptr(out), _stream(self.device)))
        return out

    # update_gradient_mixed (PS mod.rs:359-427); signs must be distinct within one call
    def update(self, signs, grads):
        signs = _as_i64_bits(signs)
        grads = grads.contiguous()
        assert grads.dtype == torch.float32 and grads.numel() == signs.numel() * self.dim
        N.check(self.lib.pb_update(self.h, _ptr(signs), _ptr(grads), signs.numel(), _stream(self.device)))

    # set_embedding (PS mod.rs:287-306)
    def set_entries(self, signs, entries):
        signs = _as_i64_bits(signs)
        entries = entries.contiguous()
        assert entries.dtype == torch.float32 and entries.numel() == signs.numel() * self.entry_len
        N.check(self.lib.pb_set_rows(self.h, _ptr(signs), _ptr(entries), signs.numel(), _stream(self.device)))

    def get_entries(self, signs):
        signs = _as_i64_bits(signs)
        n = signs.numel()
        ent = torch.empty((n, self.entry_len), dtype=torch.float32, device=self.device)
        found = torch.empty(n, dtype=torch.uint8, device=self.device)
        N.check(self.lib.pb_get_rows(self.h, _ptr(signs), n, _ptr(ent), _ptr(found), _stream(self.device)))
        return ent, found.bool()

    def spill(self, want_free, keep_batches=1, max_n=None):
        """pb_table_spill: release least recently used rows until `want_free` are free, returning what was released as
        (signs uint64 numpy, entries float32 numpy [n, entry_len]) — the host tier's input."""
        max_n = int(max_n if max_n is not None else min(self.capacity, max(int(want_free), 1)))
        signs = torch.empty(max_n, dtype=torch.int64, device=self.device)
        ent = torch.empty((max_n, self.entry_len), dtype=torch.float32, device=self.device)
        count = torch.zeros(1, dtype=torch.int32, device=self.device)
        N.check(self.lib.pb_table_spill(self.h, int(want_free), int(keep_batches), _ptr(signs), _ptr(ent), max_n, _ptr(count),
                                        _stream(self.device)))
        n = min(int(count), max_n)
        return signs[:n].cpu().numpy().view(np.uint64), ent[:n].cpu().numpy()

    def export_signs(self):
        """Resident signs (int64 bit patterns) and the training-request number each was last used in, on the device,
        sorted oldest first (ties by sign): the order the reference's LRU list is dumped in."""
        count = torch.zeros(1, dtype=torch.int32, device=self.device)
        N.check(self.lib.pb_table_export_signs(self.h, None, None, 0, _ptr(count), _stream(self.device)))
        n = int(count)
        signs = torch.empty(n, dtype=torch.int64, device=self.device)
        rec = torch.empty(n, dtype=torch.int32, device=self.device)
        if n:
            N.check(self.lib.pb_table_export_signs(self.h, _ptr(signs), _ptr(rec), n, _ptr(count), _stream(self.device)))
            assert int(count) == n, "the table changed during the export"
            order = torch.argsort(signs, stable=True)           # deterministic file: by sign ...
            order = order[torch.argsort(rec[order].to(torch.int64) & 0xFFFFFFFF, stable=True)]  # ... inside a request
            signs, rec = signs[order], rec[order]
        return signs, rec

    def close(self):
        if getattr(self, "h", None):
            self.lib.pb_table_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchContext:
    """Device-side per-batch context + workspace for the batched forward/backward of one shard."""

    def __init__(self, max_occurrences, max_out_rows, prefixes, sqrt_scaling=None, prefix_bit=8, device=0):
        self.lib = N.load()
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        self.n_slots = len(prefixes)
        self.prefixes, self.prefix_bit = [int(p) for p in prefixes], int(prefix_bit)
        h = C.c_void_p()
        N.check(self.lib.pb_ctx_create(self.device.index or 0, int(max_occurrences), int(max_out_rows), C.byref(h)))
        self.h = h
        cfg = N.SlotsCfg()
        cfg.n_slots, cfg.prefix_bit = self.n_slots, prefix_bit
        for i, p in enumerate(prefixes):
            cfg.prefix[i] = int(p)
            cfg.sqrt_scaling[i] = int(bool(sqrt_scaling[i])) if sqrt_scaling is not None else 0
        N.check(self.lib.pb_ctx_set_slots(self.h, C.byref(cfg)))

    def batch_stats(self):
        """Dedup statistics of the last training batch (host sync): distinct items and their multiplicity classes."""
        out = (C.c_uint32 * 6)()
        N.check(self.lib.pb_ctx_batch_stats(self.h, C.byref(out), _stream(self.device)))
        return {"items": out[0], "cold": out[1], "warm": out[2], "hot": out[3], "repeated_occurrences": out[4],
                "occurrences": out[5]}

    def forward(self, shard, ids, slot_occ_off, batch, row_off=None, training=True, out=None):
        """ids: flat device int64-bit ids (slot-major); slot_occ_off: host list, n_slots+1;
        row_off: device int32 CSR offsets [n_slots*batch+1] or None (one id per sample per slot).
        Returns f16 [n_slots, batch, dim]."""
        ids = _as_i64_bits(ids)
        if out is None:
            out = torch.empty((self.n_slots, batch, shard.dim), dtype=torch.float16, device=self.device)
        off = (C.c_uint32 * (self.n_slots + 1))(*[int(x) for x in slot_occ_off])
        if row_off is not None:
            assert row_off.dtype == torch.int32 and row_off.is_contiguous()
        if getattr(shard, "tier", None) is not None:  # host-DRAM tier: make room, bring the batch's spilled signs back
            shard.tier.before_lookup(add_prefix(ids, slot_occ_off, self.prefixes, self.prefix_bit))
        N.check(self.lib.pb_forward(shard.h, self.h, _ptr(ids), ids.numel(), _ptr(row_off), off, int(batch),
                                    int(training), _ptr(out), _stream(self.device)))
        return out

    def backward(self, shard, grads, scales=None, want_status=False):
        """grads: list (per slot) of device tensors [batch, dim] (all f16 or all f32) or None (skipped)."""
        ptrs = (C.c_void_p * self.n_slots)()
        is_f16 = None
        for i, g in enumerate(grads):
            if g is None:
                ptrs[i] = None
                continue
            assert g.is_contiguous()
            f16 = g.dtype == torch.float16
            assert f16 or g.dtype == torch.float32
            assert is_f16 is None or is_f16 == f16, "all slot gradients of one request share a dtype"
            is_f16 = f16
            ptrs[i] = g.data_ptr()
        sc = None
        if scales is not None:
            sc = (C.c_float * self.n_slots)(*[float(s) for s in scales])
        status = torch.empty(self.n_slots, dtype=torch.int32, device=self.device) if want_status else None
        N.check(self.lib.pb_backward(shard.h, self.h, ptrs, int(bool(is_f16)), sc, _ptr(status), _stream(self.device)))
        return status

    def backward_ptrs(self, shard, ptrs, is_f16, scales=None):
        """GradientBatch form (persia-core/src/backward.rs:86-105): raw device pointers, None = skipped slot."""
        arr = (C.c_void_p * self.n_slots)(*[(int(p) if p else None) for p in ptrs])
        sc = (C.c_float * self.n_slots)(*[float(x) for x in scales]) if scales is not None else None
        N.check(self.lib.pb_backward(shard.h, self.h, arr, int(bool(is_f16)), sc, None, _stream(self.device)))

    def forward_raw(self, shard, ids, batch, sample_fixed_size, row_off=None, training=True):
        """Raw (embedding_summation: false) slot; the context must have been built with ONE prefix.
        Returns (table f16 [n_occ+1, dim] of which rows [0, U] are valid, index i64 [batch*fixed],
        non_empty i64 [batch*fixed] of which counts[1] entries are valid, sample_id_num i32 [batch],
        counts i32 [2] = (U, non-empty entries)), all on the device (mod.rs:586-623, forward.rs:336-347)."""
        assert self.n_slots == 1, "a raw context serves one slot"
        ids = _as_i64_bits(ids)
        n, fixed = ids.numel(), int(sample_fixed_size)
        table = torch.empty((n + 1, shard.dim), dtype=torch.float16, device=self.device)
        index = torch.empty(batch * fixed, dtype=torch.int64, device=self.device)
        non_empty = torch.empty(max(batch * fixed, 1), dtype=torch.int64, device=self.device)
        sample_id_num = torch.empty(max(batch, 1), dtype=torch.int32, device=self.device)
        counts = torch.empty(2, dtype=torch.int32, device=self.device)
        if row_off is not None:
            assert row_off.dtype == torch.int32 and row_off.is_contiguous() and row_off.numel() == batch + 1
        N.check(self.lib.pb_forward_raw(shard.h, self.h, _ptr(ids), n, _ptr(row_off), int(batch), fixed, int(training),
                                        _ptr(table), _ptr(index), _ptr(non_empty), _ptr(sample_id_num), _ptr(counts),
                                        _stream(self.device)))
        return table, index, non_empty, sample_id_num[:batch], counts

    def backward_raw(self, shard, grad, scale=1.0, want_status=False, is_f16=None):
        """grad: device tensor [U, dim] (f32 as persia/ctx.py:970-980 builds it, or f16), a raw device pointer
        (then pass is_f16), or None (add_skipped_gradient)."""
        ptr = None
        if grad is not None and not isinstance(grad, int):
            assert grad.is_contiguous() and grad.dtype in (torch.float16, torch.float32)
            is_f16 = grad.dtype == torch.float16
            ptr = grad.data_ptr() if grad.numel() else self._raw_dummy().data_ptr()
        elif grad is not None:
            ptr = grad
        status = torch.empty(1, dtype=torch.int32, device=self.device) if want_status else None
        N.check(self.lib.pb_backward_raw(shard.h, self.h, ptr, int(bool(is_f16)), float(scale), _ptr(status),
                                         _stream(self.device)))
        return status

    def _raw_dummy(self):
        if not hasattr(self, "_dummy"):
            self._dummy = torch.zeros(4, dtype=torch.float32, device=self.device)
        return self._dummy

    def close(self):
        if getattr(self, "h", None):
            self.lib.pb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def add_prefix(ids, slot_occ_off, prefixes, prefix_bit=8):
    lib = N.load()
    ids = _as_i64_bits(ids)
    out = torch.empty_like(ids)
    S = len(prefixes)
    off = (C.c_uint32 * (S + 1))(*[int(x) for x in slot_occ_off])
    pf = (C.c_uint64 * S)(*[int(p) for p in prefixes])
    N.check(lib.pb_add_prefix(_ptr(ids), ids.numel(), off, pf, S, prefix_bit, _ptr(out), _stream(ids.device)))
    return out


def shard_of(signs, R):
    lib = N.load()
    signs = _as_i64_bits(signs)
    out = torch.empty(signs.numel(), dtype=torch.int32, device=signs.device)
    N.check(lib.pb_shard_of(_ptr(signs), signs.numel(), R, _ptr(out), _stream(signs.device)))
    return out


def hash_stack(ids, rounds, embedding_size, out=None):
    """indices_to_hashstack_indices (mod.rs:347-400) on the device: [n] ids -> [n * rounds] keys, id-major."""
    lib = N.load()
    ids = _as_i64_bits(ids)
    if out is None:
        out = torch.empty(ids.numel() * int(rounds), dtype=torch.int64, device=ids.device)
    N.check(lib.pb_hash_stack(_ptr(ids), ids.numel(), int(rounds), int(embedding_size), _ptr(out), _stream(ids.device)))
    return out


def farmhash64(x):
    lib = N.load()
    x = _as_i64_bits(x)
    out = torch.empty_like(x)
    N.check(lib.pb_farmhash64(_ptr(x), x.numel(), _ptr(out), _stream(x.device)))
    return out


def partition_by_shard(signs, R):
    """Stable partition: returns (perm int32[n], counts int32[R])."""
    lib = N.load()
    signs = _as_i64_bits(signs)
    n = signs.numel()
    perm = torch.empty(n, dtype=torch.int32, device=signs.device)
    counts = torch.empty(R, dtype=torch.int32, device=signs.device)
    wb = int(lib.pb_partition_workspace(n))
    work = torch.empty(wb, dtype=torch.uint8, device=signs.device)
    N.check(lib.pb_partition_by_shard(_ptr(signs), n, R, _ptr(perm), _ptr(counts), _ptr(work), wb, _stream(signs.device)))
    return perm, counts

"""Shared helpers for the parity tests: seeded synthetic id batches in the reference's LIL shape,
flattened to the slot-major CSR the C ABI takes."""
import numpy as np


def make_batch(rng, n_slots, batch, cardinality, max_ids=1, allow_empty=False):
    """Returns (ids u64 [N], row_off u32 [S*B+1] or None, slot_occ_off list[S+1]).
    max_ids == 1 and not allow_empty -> single-id slots (row_off None)."""
    card = [cardinality] * n_slots if np.isscalar(cardinality) else list(cardinality)
    if max_ids == 1 and not allow_empty:
        ids = np.stack([rng.integers(0, card[s], size=batch, dtype=np.uint64) for s in range(n_slots)]).reshape(-1)
        slot_off = [s * batch for s in range(n_slots + 1)]
        return ids, None, slot_off
    counts = rng.integers(0 if allow_empty else 1, max_ids + 1, size=n_slots * batch)
    row_off = np.zeros(n_slots * batch + 1, np.uint32)
    row_off[1:] = np.cumsum(counts)
    ids = np.empty(int(row_off[-1]), np.uint64)
    for s in range(n_slots):
        lo, hi = int(row_off[s * batch]), int(row_off[(s + 1) * batch])
        ids[lo:hi] = rng.integers(0, card[s], size=hi - lo, dtype=np.uint64)
    slot_off = [int(row_off[s * batch]) for s in range(n_slots + 1)]
    return ids, row_off, slot_off


def full_row_off(n_slots, batch):
    return np.arange(n_slots * batch + 1, dtype=np.uint32)


def to_dev_ids(ids, device):
    import torch

    return torch.from_numpy(np.ascontiguousarray(ids).view(np.int64)).to(device)


def to_dev_i32(a, device):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(device)


def f16_ulp_diff(a, b):
    """Max distance in f16 representable steps between two float16 arrays (sign-magnitude -> ordered ints)."""
    def key(x):
        u = x.view(np.uint16).astype(np.int32)
        return np.where(u & 0x8000, -(u & 0x7FFF), u & 0x7FFF)
    return int(np.max(np.abs(key(np.ascontiguousarray(a)) - key(np.ascontiguousarray(b))))) if a.size else 0

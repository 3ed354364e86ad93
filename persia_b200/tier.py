"""Host-DRAM tier behind a GPU-resident shard (BASELINE.json configs[4]; SURVEY §8f "next").

The reference's embedding holder keeps every entry in host DRAM (persia-embedding-holder); here a shard's HBM holds
the working set and this class holds the rest: rows the shard releases are written out first (pb_table_spill: sign +
whole entry, embedding ++ optimizer state) and kept here; before a lookup the batch's signs that are not resident are
looked up here and put back with pb_set_rows (the reference's set_embedding).  A sign is in exactly one place, so the
shard + the tier behave like one table of unbounded capacity — `tests/test_gpu_tier.py` checks that against the
oracle bit for bit.  This is the functional tier (sorted runs in host memory, synchronous staging); the asynchronous
staging stream and a pinned open-addressing store are the next step.
"""
import numpy as np
import torch


class _Runs:
    """sign -> entry store in host memory: a few sorted runs (keys uint64, values float32 [n, L]); lookups are vectorised
    binary searches, a spill appends a run, runs are merged when there are too many."""

    def __init__(self, max_runs=64):
        self.runs, self.max_runs, self.n = [], max_runs, 0

    def __len__(self):
        return self.n

    def add(self, keys, vals):
        if not keys.size:
            return
        o = np.argsort(keys, kind="stable")
        self.runs.append([keys[o], vals[o], np.ones(keys.size, bool)])
        self.n += int(keys.size)
        if len(self.runs) > self.max_runs:
            k = np.concatenate([r[0][r[2]] for r in self.runs])
            v = np.concatenate([r[1][r[2]] for r in self.runs])
            o = np.argsort(k, kind="stable")
            self.runs = [[k[o], v[o], np.ones(k.size, bool)]]

    def take(self, keys):
        """Entries of the keys present (removed from the store): (found keys, entries)."""
        got_k, got_v = [], []
        for r in self.runs:
            if not keys.size:
                break
            i = np.searchsorted(r[0], keys)
            i[i >= r[0].size] = 0
            ok = (r[0][i] == keys) & r[2][i] if r[0].size else np.zeros(keys.size, bool)
            if ok.any():
                got_k.append(keys[ok])
                got_v.append(r[1][i[ok]])
                r[2][i[ok]] = False
                keys = keys[~ok]
        if not got_k:
            return np.zeros(0, np.uint64), None
        k = np.concatenate(got_k)
        self.n -= int(k.size)
        return k, np.concatenate(got_v)

    def get(self, key):
        for r in self.runs:
            i = int(np.searchsorted(r[0], np.uint64(key)))
            if i < r[0].size and r[0][i] == np.uint64(key) and r[2][i]:
                return r[1][i]
        return None


class HostTier:
    def __init__(self, shard, reserve, keep_batches=1):
        """reserve: rows that must be free before a lookup (an upper bound on the distinct signs of a batch)."""
        self.shard, self.reserve, self.keep = shard, int(reserve), int(keep_batches)
        self.store = _Runs()
        self.spilled = self.restored = self.lookups = self.hits_gpu = 0
        shard.tier = self

    def __len__(self):
        return len(self.store)

    def before_lookup(self, signs_dev):
        sh = self.shard
        if sh.capacity - len(sh) < self.reserve:  # make room first: what this batch still needs comes back below
            s, e = sh.spill(self.reserve, keep_batches=self.keep, max_n=max(self.reserve, 1024))
            self.store.add(s, e)
            self.spilled += int(s.size)
            if sh.capacity - len(sh) < self.reserve:
                raise RuntimeError("host tier: the shard cannot free the reserve (rows of in-flight batches are protected): "
                                   "raise the shard's capacity or lower `reserve`")
        if not len(self.store):
            return
        u = torch.unique(signs_dev)
        _, found = sh.get_entries(u)
        self.lookups += int(u.numel())
        self.hits_gpu += int(found.sum())
        miss = u[~found].cpu().numpy().view(np.uint64)
        back, ent = self.store.take(miss)
        if back.size:
            sh.set_entries(torch.from_numpy(back.view(np.int64)).to(sh.device), torch.from_numpy(ent).to(sh.device))
            self.restored += int(back.size)

    def get_entry(self, sign):
        """An entry wherever it lives (tests, checkpoints)."""
        e = self.store.get(sign)
        if e is not None:
            return e
        ent, found = self.shard.get_entries(torch.from_numpy(np.array([sign], np.uint64).view(np.int64)).to(self.shard.device))
        return ent[0].cpu().numpy() if bool(found[0]) else None

    def stats(self):
        return {"host_rows": len(self.store), "spilled": self.spilled, "restored": self.restored,
                "gpu_hit_ratio": (self.hits_gpu / self.lookups) if self.lookups else None}

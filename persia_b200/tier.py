"""Host-DRAM tier behind a GPU-resident shard (BASELINE.json configs[4]; SURVEY §8f "next").

The reference's embedding holder keeps every entry in host DRAM (persia-embedding-holder); here a shard's HBM holds
the working set and this class holds the rest: rows the shard releases are written out first (pb_table_spill: sign +
whole entry, embedding ++ optimizer state) and kept here; before a lookup the batch's signs that are not resident are
looked up here and put back with pb_set_rows (the reference's set_embedding).  A sign is in exactly one place, so the
shard + the tier behave like one table of unbounded capacity — `tests/test_gpu_tier.py` checks that against the
oracle bit for bit.  This is the functional tier (host dictionary, synchronous staging); the asynchronous staging
stream and a pinned open-addressing store are the next step.
"""
import numpy as np
import torch


class HostTier:
    def __init__(self, shard, reserve, keep_batches=1):
        """reserve: rows that must be free before a lookup (an upper bound on the distinct signs of a batch)."""
        self.shard, self.reserve, self.keep = shard, int(reserve), int(keep_batches)
        self.store = {}  # sign -> float32 entry
        self.spilled = self.restored = self.lookups = self.hits_gpu = 0
        shard.tier = self

    def __len__(self):
        return len(self.store)

    def before_lookup(self, signs_dev):
        sh = self.shard
        free = sh.capacity - len(sh)
        if free < self.reserve:  # make room first: what this batch still needs comes back below
            s, e = sh.spill(self.reserve, keep_batches=self.keep, max_n=max(self.reserve, 1024))
            for k in range(s.size):
                self.store[int(s[k])] = e[k].copy()
            self.spilled += int(s.size)
            if sh.capacity - len(sh) < self.reserve:
                raise RuntimeError("host tier: the shard cannot free the reserve (rows of in-flight batches are protected): "
                                   "raise the shard's capacity or lower `reserve`")
        if not self.store:
            return
        u = torch.unique(signs_dev)
        _, found = sh.get_entries(u)
        self.lookups += int(u.numel())
        self.hits_gpu += int(found.sum())
        miss = u[~found].cpu().numpy().view(np.uint64)
        back = [int(x) for x in miss if int(x) in self.store]
        if back:
            ent = np.stack([self.store.pop(x) for x in back])
            sh.set_entries(torch.from_numpy(np.array(back, np.uint64).view(np.int64)).to(sh.device),
                           torch.from_numpy(ent).to(sh.device))
            self.restored += len(back)

    def get_entry(self, sign):
        """An entry wherever it lives (tests, checkpoints)."""
        if int(sign) in self.store:
            return self.store[int(sign)]
        ent, found = self.shard.get_entries(torch.tensor([np.uint64(sign).astype(np.int64)], device=self.shard.device))
        return ent[0].cpu().numpy() if bool(found[0]) else None

    def stats(self):
        return {"host_rows": len(self.store), "spilled": self.spilled, "restored": self.restored,
                "gpu_hit_ratio": (self.hits_gpu / self.lookups) if self.lookups else None}

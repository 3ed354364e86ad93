"""configs[4]-shaped measurement at reduced scale: a shard whose HBM holds a fraction of the key space + the host tier.
Reports samples/s of training steps through the tier, the GPU hit ratio and the rows staged per step."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from persia_b200 import native as N
from persia_b200 import shard as SH
from persia_b200 import workload as W
from persia_b200.tier import HostTier

dim, B, S = 96, 4096, 26
cap, keys = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_500_000, int(float(sys.argv[2])) if len(sys.argv) > 2 else 20_000_000
steps = 60
dev = torch.device("cuda", 0)
card = W.scaled_cardinalities(keys, S)
pf = W.index_prefixes(S)
sh = SH.EmbeddingShard(dim, cap, dev)
sh.set_optimizer(N.OPT_ADAGRAD, lr=0.01, initialization=0.01, eps=1e-10)
sh.configure()
tier = HostTier(sh, reserve=S * B)
ctx = SH.BatchContext(S * B, S * B, pf, device=dev)
ids = W.make_batches(5, card, B, steps + 40, 1.05)
grads = [(torch.randn((B, dim), device=dev) * 1e-2).half() for _ in range(S)]
slot_off = [s * B for s in range(S + 1)]
out = torch.empty((S, B, dim), dtype=torch.float16, device=dev)


def step(k):
    ctx.forward(sh, torch.from_numpy(ids[k].view(np.int64)).to(dev), slot_off, B, training=True, out=out)
    ctx.backward(sh, grads)


WARM = 40
for k in range(WARM):  # fill the shard and start spilling
    step(k)
torch.cuda.synchronize()
s0 = dict(tier.stats())
t0 = time.time()
for k in range(WARM, WARM + steps):
    step(k)
torch.cuda.synchronize()
dt = time.time() - t0
s1 = tier.stats()
print(json.dumps({"workload": f"configs[4] shape at reduced scale: 26 slots, dim {dim}, batch {B}, {keys:.3g}-id Zipf key space, {cap:.3g} rows in HBM",
                  "samples_per_s": B * steps / dt, "ms_per_step": 1e3 * dt / steps, "gpu_resident_rows": len(sh),
                  "host_rows": s1["host_rows"], "gpu_hit_ratio_of_distinct_signs": s1["gpu_hit_ratio"],
                  "rows_spilled_per_step": (s1["spilled"] - s0["spilled"]) / steps,
                  "rows_restored_per_step": (s1["restored"] - s0["restored"]) / steps,
                  "capacity_refused": sh.counters()["capacity_refused"],
                  "note": "functional tier: host store = sorted numpy runs, staging synchronous on the lookup's stream"}))

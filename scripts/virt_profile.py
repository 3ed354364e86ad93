"""R virtual ranks on ONE GPU at the bench's per-rank sizes, for an ncu launch list of the sharded step's kernels
(ncu replays each kernel and cannot follow R processes; the kernels of one rank are the same launches either way)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from persia_b200 import native as N
from persia_b200 import workload as W
from persia_b200.worker import ShardedEmbeddingWorker as SW

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dim, B, S = 128, 8192, 26
dev = torch.device("cuda", 0)
card = W.scaled_cardinalities(int(1e8) * R, S)
pf = W.index_prefixes(S)
ids = [W.make_batches(7 + r, card, B, 2, 1.05) for r in range(R)]
cap = SW.calibrate_cap([ids[r][k] for r in range(R) for k in range(2)], B, pf, R)
ws = SW.local_group(R, S, dim, pf, 1 << 21, cap=cap, optimizer=dict(kind=N.OPT_ADAGRAD, lr=0.01, initialization=0.01, eps=1e-10),
                    max_batch=B)
for x in ws:
    x.shard.get_entries(torch.zeros(1, dtype=torch.int64, device=dev))
torch.cuda.synchronize()
outs = [torch.empty((S, B, dim), dtype=torch.float16, device=dev) for _ in range(R)]
grads = [[(torch.randn((B, dim), device=dev) * 1e-2).half() for _ in range(S)] for _ in range(R)]
for step in range(6):
    d_ids = [torch.from_numpy(ids[r][step % 2].view(np.int64)).to(dev) for r in range(R)]
    torch.cuda.synchronize()
    SW.group_forward(ws, d_ids, B, training=True, outs=outs)
    torch.cuda.synchronize()
    SW.group_backward(ws, grads)
    torch.cuda.synchronize()
print("ok", R, "cap", cap, [x.status() for x in ws][:2])

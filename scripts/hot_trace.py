"""Debugging aid: per hot item timestamps of k_reduce_hot (start, order phase done, sum done) at configs[1] shape."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from persia_b200 import native as N
from persia_b200 import shard as SH
from persia_b200 import workload as W

dim, B, S, rows = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 4096, 26, 10_000_000
lib = N.load()
dev = torch.device("cuda", 0)
card = W.scaled_cardinalities(int(1e8), S)
pf = W.index_prefixes(S)
sh = SH.EmbeddingShard(dim, rows, dev)
sh.set_optimizer(N.OPT_ADAGRAD, lr=0.01, initialization=0.01, eps=1e-10)
sh.configure()
n_occ = S * B
ctx = SH.BatchContext(n_occ, n_occ, pf, device=dev)
ids = W.make_batches(2, card, B, 4, 1.05)
ids_dev = [torch.from_numpy(ids[k].view(np.int64)).to(dev) for k in range(4)]
grads = [[(torch.randn((B, dim), device=dev) * 1e-2).half() for _ in range(S)] for _ in range(4)]
out = torch.empty((S, B, dim), dtype=torch.float16, device=dev)
slot_off = [s * B for s in range(S + 1)]
trace = torch.zeros(8 * 8192, dtype=torch.int64, device=dev)
for k in range(6):
    ctx.forward(sh, ids_dev[k % 4], slot_off, B, training=True, out=out)
    ctx.backward(sh, grads[k % 4])
torch.cuda.synchronize()
lib.pb_debug_hot_trace.argtypes = [C.c_void_p]
lib.pb_debug_hot_trace(C.c_void_p(trace.data_ptr()))
ctx.forward(sh, ids_dev[2], slot_off, B, training=True, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ctx.backward(sh, grads[2])
e1.record()
torch.cuda.synchronize()
lib.pb_debug_hot_trace(C.c_void_p(0))
t = trace.cpu().numpy().reshape(-1, 8)
t = t[t[:, 0] != 0]
t0 = t[:, 0].min()
cnt = t[:, 3] & 0xFFFFFFFF
blk = t[:, 3] >> 32
print("backward ms", e0.elapsed_time(e1), "hot items", len(t), "span us", (t[:, 2].max() - t0) / 1e3)
order = np.argsort(-cnt)
print("  cnt  blk  start_us  order_us  sum_us  ns/row | chain: wait_cyc add_cyc | producer 0: wait_cyc work_cyc (per its chunks)")
for i in list(order[:12]) + list(order[-4:]):
    print("%5d %4d %8.2f %8.2f %8.2f %7.1f | %8d %8d | %8d %8d" % (
        cnt[i], blk[i], (t[i, 0] - t0) / 1e3, (t[i, 1] - t[i, 0]) / 1e3, (t[i, 2] - t[i, 1]) / 1e3, (t[i, 2] - t[i, 1]) / max(1, cnt[i]),
        t[i, 4], t[i, 5], t[i, 6], t[i, 7]))
late = np.argsort(-t[:, 2])[:6]
print("last to finish:")
for i in late:
    print("%5d %4d start %8.2f end %8.2f" % (cnt[i], blk[i], (t[i, 0] - t0) / 1e3, (t[i, 2] - t0) / 1e3))

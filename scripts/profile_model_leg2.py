"""torch.profiler view of the e2e_model leg: where the GPU time of a TrainCtx step goes."""
import sys

sys.path.insert(0, ".")
import torch
from torch.profiler import ProfilerActivity, profile

import bench

args = bench.parse_args()
args.rows = 2e7
orig = bench.model_leg


def run():
    return orig(args, torch, steps=20, warmup=10)


with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    res = run()
print(res["value"], res["ms_per_step"])
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))

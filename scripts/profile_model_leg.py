"""cProfile of bench.py's e2e_model leg (host side of the TrainCtx path)."""
import cProfile
import pstats
import sys

sys.path.insert(0, ".")
import torch

import bench

args = bench.parse_args()
args.rows = 2e7
pr = cProfile.Profile()
pr.enable()
res = bench.model_leg(args, torch, steps=40, warmup=10)
pr.disable()
print(res["value"], res["ms_per_step"])
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_callees("api.py:.*\\((backward|_on_backward|get_embedding_from_data|forward)\\)")
st.sort_stats("tottime").print_stats(25)

#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
n=8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench$n.json 2> gpurun_out/r2s_bench$n.err
echo "bench $n exit $?"; tail -3 gpurun_out/r2s_bench$n.err | cut -c1-300; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2s_bench$n.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","parity_checked") if k in d}, "e2e", d.get("e2e",{}).get("value"))
    print("model", d.get("e2e_model"))
except Exception as e: print("no json",e)
PY

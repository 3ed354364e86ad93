#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
n=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2w_bench$n.json 2> gpurun_out/r2w_bench$n.err
echo "bench $n exit $?"; tail -3 gpurun_out/r2w_bench$n.err | cut -c1-300; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2w_bench$n.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","parity_checked") if k in d}, "e2e", d.get("e2e",{}).get("value"))
except Exception as e: print("no json",e)
PY
timeout 600 python -m pytest tests/test_gpu_worker.py tests/test_gpu_trainctx.py -m gpu -q -p no:cacheprovider -x -k "two_process" 2>&1 | tail -3

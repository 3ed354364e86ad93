#!/bin/bash
# two GPUs: the 2-process tests, then the N=2 bench through torchrun
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_worker.py tests/test_gpu_trainctx.py -m gpu -q -p no:cacheprovider -x -k "two_process" > gpurun_out/r2q_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2q_pytest.log
grep -E "passed|failed|FAILED|pytest exit|Error|error" gpurun_out/r2q_pytest.log | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2q_bench2.json 2> gpurun_out/r2q_bench2.err
echo "bench exit $?"; tail -5 gpurun_out/r2q_bench2.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2q_bench2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","parity_checked") if k in d}, "e2e", d.get("e2e"))
    print(json.dumps(d.get("run"))[:1500])
except Exception as e: print("no json",e)
PY

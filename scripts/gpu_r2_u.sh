#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2u_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2u_pytest.log
grep -E "passed|failed|FAILED|pytest exit|Mismatched|Error" gpurun_out/r2u_pytest.log | tail -12
bash scripts/gpu_r2_t.sh 2>&1 | tail -18

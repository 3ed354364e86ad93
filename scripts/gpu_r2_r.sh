#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests/test_gpu_trainctx.py -m gpu -q -p no:cacheprovider -x -k "two_process" > gpurun_out/r2r_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2r_pytest.log
grep -E "passed|failed|FAILED|pytest exit|Error|error" gpurun_out/r2r_pytest.log | tail -15

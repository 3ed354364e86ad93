#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2n_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
grep -E "passed|failed|FAILED|pytest exit|Mismatched|Error" gpurun_out/r2n_pytest.log | tail -12
timeout 300 python scripts/hot_trace.py 64 4096 > gpurun_out/r2n_trace64.txt 2>&1; cat gpurun_out/r2n_trace64.txt | tail -26
timeout 300 python scripts/hot_trace.py 128 8192 > gpurun_out/r2n_trace128.txt 2>&1; cat gpurun_out/r2n_trace128.txt | tail -26

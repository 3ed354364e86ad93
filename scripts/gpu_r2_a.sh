#!/bin/bash
# round 2, first GPU pass: parity tests of the single-GPU path, a short bench, a launch list
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_worker.py -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench exit $?"; tail -5 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json

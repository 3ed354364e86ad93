#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2e_pytest.log
grep -E "passed|failed|FAILED|pytest exit|Mismatched" gpurun_out/r2e_pytest.log | tail -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_reduce_hot|k_probe_items|k_reduce_items|k_dedup|k_gather_items" -s 10 -c 5 -o gpurun_out/r2e_kern -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity > gpurun_out/r2e_full.log 2>&1
ls -la gpurun_out/r2e_kern.ncu-rep

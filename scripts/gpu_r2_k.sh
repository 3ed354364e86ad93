#!/bin/bash
# ncu --set full capture of one launch of each kernel of the step at configs[1] (dim 64, batch 4096)
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'k_reduce_warm|k_reduce_cold|k_reduce_hot|k_dedup|k_probe_items|k_gather_items|k_nan_scan|k_clear_items' \
  --launch-skip 160 -c 8 -f -o gpurun_out/r2k_full \
  python bench.py --steps 20 --warmup 10 --batch 4096 --dim 64 --no-cpu-baseline --no-graph --no-parity --no-model-leg --no-roofline-leg > gpurun_out/r2k_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/r2k_ncu.log
ls -la gpurun_out/r2k_full.ncu-rep

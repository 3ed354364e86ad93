#!/bin/bash
# round-2 evidence: launch list of the bench command (both legs) and one --set full capture per kernel of the step at
# configs[1]; summarised into profiles/ by scripts/r2_profile_summary.py (run where ncu is, on the merged outputs)
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity --no-model-leg --no-staleness"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_prof1.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_reduce_warm|k_reduce_cold|k_reduce_hot|k_dedup|k_probe_items|k_gather_items|k_nan_scan|k_clear_items' \
  --launch-skip 160 -c 8 -f -o gpurun_out/r2_full \
  python bench.py --steps 20 --warmup 10 --batch 4096 --dim 64 --no-cpu-baseline --no-graph --no-parity --no-model-leg --no-roofline-leg --no-staleness > gpurun_out/r2_prof2.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out/r2_full.ncu-rep

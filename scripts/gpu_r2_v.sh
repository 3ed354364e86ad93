#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_owner_update_all' --launch-skip 36 -c 1 -f -o gpurun_out/r2v_owner python scripts/virt_profile.py 8 > gpurun_out/r2v.log 2>&1
tail -2 gpurun_out/r2v.log; ls -la gpurun_out/r2v_owner.ncu-rep

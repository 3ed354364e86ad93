#!/bin/bash
# hot-kernel diagnosis: plain-load variant vs bulk ring, and a full ncu capture of the kernel
mkdir -p gpurun_out
PB_HOT_NO_BULK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_nobulk.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2c_nobulk.log 2>&1
python profiles/launch_list.py gpurun_out/r2c_nobulk.csv 10
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_reduce_hot -s 6 -c 1 -o gpurun_out/r2c_hot -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2c_full.log 2>&1
ls -la gpurun_out/r2c_hot.ncu-rep

#!/bin/bash
# round 2, second GPU pass: whole parity suite incl. the sharded path on virtual ranks, bench, ncu launch list
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -40 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench exit $?"; tail -5 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r2b_ncu.log 2>&1
python profiles/launch_list.py gpurun_out/r2b_launches.csv 2>/dev/null | tail -40

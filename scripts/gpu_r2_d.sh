#!/bin/bash
# round 2: whole parity suite (sharded path on phase-driven virtual ranks), new bench (two legs + parity), launch list
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
grep -E "passed|failed|Error|FAILED|pytest exit" gpurun_out/r2d_pytest.log | tail -30
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench exit $?"; tail -5 gpurun_out/r2d_bench.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2d_bench.json'))
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"parity",d["parity_checked"])
    print("roof leg",d["run"]["roofline_leg"])
    r=d["roofline"]; print("frac",r["frac"],"worst",r["frac_worst_case"]); 
    for k,v in r["kernels"].items(): print(" ",k,v)
    print(" whole",r["whole_step"])
    for k,v in r["metric_leg"]["kernels"].items(): print(" m",k,v)
    print(" m whole",r["metric_leg"]["whole_step"])
except Exception as e: print("no json",e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity > gpurun_out/r2d_ncu.log 2>&1
grep -E "k_reduce|k_dedup|k_probe_items|k_gather_items|k_nan|k_clear|k_begin" gpurun_out/r2d_launches.csv | awk -F'","' '{print $5, $(NF)}' | tr -d '"' | sed 's/(.*)//' | awk '{n[$1" "$2" "$3]++; s[$1" "$2" "$3]+=$NF} END {for (k in n) printf "%s n=%d avg=%.2f us\n", k, n[k], s[k]/n[k]/1000}' | sort | head -40

#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2p_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2p_pytest.log
grep -E "passed|failed|FAILED|pytest exit|Mismatched|Error" gpurun_out/r2p_pytest.log | tail -12
timeout 300 python scripts/hot_trace.py 64 4096 > gpurun_out/r2p_trace64.txt 2>&1; cat gpurun_out/r2p_trace64.txt | tail -26 | cut -c1-80
timeout 300 python scripts/hot_trace.py 128 8192 > gpurun_out/r2p_trace128.txt 2>&1; cat gpurun_out/r2p_trace128.txt | tail -26 | cut -c1-80
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-model-leg > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
echo "bench exit $?"; tail -5 gpurun_out/r2p_bench.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2p_bench.json'))
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"parity",d["parity_checked"])
    rl=d["run"]["roofline_leg"]; print("roof leg ms",rl["ms_per_step"],"samples/s",rl["samples_per_s"],"parity",rl["parity"] and rl["parity"]["checked"])
    print("in flight", d["run"]["batches_in_flight"])
    r=d["roofline"]; print("frac",r["frac"],"worst",r["frac_worst_case"]); 
    for k,v in r["kernels"].items(): print(" ",k,v)
    for k,v in r["metric_leg"]["kernels"].items(): print(" m",k,v)
except Exception as e: print("no json",e)
PY

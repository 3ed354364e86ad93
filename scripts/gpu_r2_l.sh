#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2l_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2l_pytest.log
grep -E "passed|failed|FAILED|pytest exit|Mismatched" gpurun_out/r2l_pytest.log | tail -20
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-model-leg > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench exit $?"; tail -5 gpurun_out/r2l_bench.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2l_bench.json'))
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"parity",d["parity_checked"])
    rl=d["run"]["roofline_leg"]; print("roof leg ms",rl["ms_per_step"],"samples/s",rl["samples_per_s"],"parity",rl["parity"] and rl["parity"]["checked"])
    r=d["roofline"]; print("frac",r["frac"],"worst",r["frac_worst_case"]); 
    for k,v in r["kernels"].items(): print(" ",k,v)
    print(" whole",r["whole_step"])
    for k,v in r["metric_leg"]["kernels"].items(): print(" m",k,v)
    print(" m whole",r["metric_leg"]["whole_step"])
except Exception as e: print("no json",e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity --no-model-leg > gpurun_out/r2l_ncu.log 2>&1
python - <<'PY'
import csv,re,collections
lines=[l for l in open('gpurun_out/r2l_launches.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict()
for r in rows:
    n=re.sub(r'\(.*','',r['Kernel Name'])
    if not any(k in n for k in ('k_reduce','k_dedup','k_probe_items','k_gather_items','k_nan','k_clear')): continue
    key=(n,r['Grid Size'],r['Block Size'])
    agg.setdefault(key,[]).append(float(r['Metric Value'])/1e3)
for k,v in agg.items(): print(k, 'n=%d'%len(v), 'avg=%.1f us min=%.1f max=%.1f'%(sum(v)/len(v),min(v),max(v)))
PY

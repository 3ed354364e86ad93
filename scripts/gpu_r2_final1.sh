#!/bin/bash
# final single-GPU evidence: the driver's own sequence (pytest -m gpu, smoke, default bench, reference arm)
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; tail -2 gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench exit $? in $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "reference arm exit $? in $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_n1.json'))
print({k:d[k] for k in ("value","ms_per_step","steps","warmup","parity_checked","gpu_launches")}, "e2e", d["e2e"]["value"], "model", d["e2e_model"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print("roofline", d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["frac_worst_case"], d["roofline"]["whole_step"]["frac"])
print("leg", d["run"]["roofline_leg"]["ms_per_step"], d["run"]["roofline_leg"]["samples_per_s"], d["run"]["batches_in_flight"]["2"], d["run"]["batches_in_flight"]["4"])
for k,v in d["roofline"]["kernels"].items(): print(" ",k,v)
r=json.load(open('gpurun_out/r2f_bench_ref.json')); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
bash scripts/r2_profile.sh 2>&1 | tail -3

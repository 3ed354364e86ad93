#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import csv,re,collections,sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict()
for r in rows:
    n=re.sub(r'\(.*','',r['Kernel Name'])
    if not any(k in n for k in ('k_reduce_hot',)): continue
    key=(n,r['Grid Size'],r['Block Size'])
    agg.setdefault(key,[]).append(float(r['Metric Value'])/1e3)
for k,v in agg.items(): print(sys.argv[1], k, 'n=%d'%len(v), 'avg=%.1f us min=%.1f max=%.1f'%(sum(v)/len(v),min(v),max(v)))
PY
}
A="--steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity --no-roofline-leg --dim 64 --batch 4096"
PB_HOT_NO_BULK=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_reduce_hot -c 20 --csv --log-file gpurun_out/r2g_nobulk.csv python bench.py $A > gpurun_out/r2g_a.log 2>&1
summ gpurun_out/r2g_nobulk.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_reduce_hot -c 20 --csv --log-file gpurun_out/r2g_bulk.csv python bench.py $A > gpurun_out/r2g_b.log 2>&1
summ gpurun_out/r2g_bulk.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_reduce_hot -s 6 -c 1 -o gpurun_out/r2g_hot -f python bench.py $A > gpurun_out/r2g_full.log 2>&1
ls -la gpurun_out/r2g_hot.ncu-rep

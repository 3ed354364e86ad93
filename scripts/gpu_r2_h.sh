#!/bin/bash
mkdir -p gpurun_out
A="--steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-parity --no-roofline-leg --dim 64 --batch 4096"
run() { # name env...
  name=$1; shift
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.max --clock-control none -k regex:k_reduce_hot -s 4 -c 6 --csv --log-file gpurun_out/r2h_$name.csv python bench.py $A > gpurun_out/r2h_$name.log 2>&1
  python - "$name" <<'PY'
import csv,sys
lines=[l for l in open('gpurun_out/r2h_%s.csv'%sys.argv[1]) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
d=[float(r['Metric Value']) for r in rows if r['Metric Name']=='gpu__time_duration.sum']
c=[float(r['Metric Value']) for r in rows if r['Metric Name']=='sm__cycles_active.max']
print(sys.argv[1], 'time us avg %.1f'%(sum(d)/len(d)/1e3), 'max cycles avg %.0f'%(sum(c)/len(c)))
PY
}
run base X=1
run noconv PB_HOT_DBG=1
run nochain PB_HOT_DBG=2
run nocopy PB_HOT_DBG=4
run none PB_HOT_DBG=7
run cs8 PB_HOT_CS=8
run rs4 PB_HOT_RS=4

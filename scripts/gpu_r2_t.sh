#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_virt8.csv python scripts/virt_profile.py 8 > gpurun_out/r2t_virt8.log 2>&1
tail -3 gpurun_out/r2t_virt8.log
python - <<'PY'
import csv,re,collections
lines=[l for l in open('gpurun_out/r2t_virt8.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
# keep the last 2 steps: find launches per step by counting k_dedup
names=[re.sub(r'\(.*','',r['Kernel Name']) for r in rows]
idx=[i for i,n in enumerate(names) if 'k_dedup' in n]
R=8
start=idx[-2*R] if len(idx)>=2*R else 0
agg=collections.OrderedDict()
for r,n in list(zip(rows,names))[start:]:
    if not n.startswith('void pb::') and not n.startswith('pb::'): continue
    key=(n,r['Grid Size'],r['Block Size'])
    agg.setdefault(key,[]).append(float(r['Metric Value'])/1e3)
tot=0
for k,v in agg.items():
    per_rank_step=sum(v)/(2*R)
    tot+=per_rank_step
    print('%-60s grid %-14s n=%3d avg=%6.1f us  per rank-step=%6.1f us'%(k[0][:60],k[1],len(v),sum(v)/len(v),per_rank_step))
print("sum per rank-step: %.1f us"%tot)
PY

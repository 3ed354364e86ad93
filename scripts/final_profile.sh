# launch lists (cold and warm caches) and a full-set capture of one step, same command as the bench leg
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/r1_launches_final_cold.csv $CMD > gpurun_out/p1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:k_ --csv --log-file gpurun_out/r1_launches_final_warm.csv $CMD > gpurun_out/p2.log 2>&1
# position of the last step of the run (its k_begin_batch is the last one in the list)
SKIP=$(python - <<'PY'
import csv
rows = list(csv.DictReader(l for l in open("gpurun_out/r1_launches_final_warm.csv") if not l.startswith("==")))
pos = [i for i, r in enumerate(rows) if "k_begin_batch" in r["Kernel Name"]]
print(pos[-1])
PY
)
echo "skip=$SKIP"
timeout 400 ncu --set full --import-source on --clock-control none --cache-control none -k regex:k_ -s $SKIP -c 12 -o gpurun_out/r1_full_final $CMD > gpurun_out/p3.log 2>&1
ls -la gpurun_out/ | tail -8

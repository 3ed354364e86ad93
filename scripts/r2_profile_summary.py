"""Summaries of the round-2 ncu outputs for profiles/ (run here after gpurun merged them back)."""
import collections
import csv
import json
import re
import subprocess
import sys

out = {}
rows = list(csv.DictReader(l for l in open("gpurun_out/r2_launches.csv") if not l.startswith("==")))
agg = collections.OrderedDict()
for r in rows:
    n = re.sub(r"\(.*", "", r["Kernel Name"])
    if not any(k in n for k in ("k_dedup", "k_probe_items", "k_gather_items", "k_clear_items", "k_nan_scan", "k_reduce_")):
        continue  # (table fill and parity helpers)
    agg.setdefault((n, r["Grid Size"], r["Block Size"]), []).append(float(r["Metric Value"]) / 1e3)
with open("profiles/r2_launch_list.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_  python bench.py --steps 3 --warmup 3 "
            "--no-cpu-baseline --no-graph --no-parity --no-model-leg --no-staleness\n"
            "# per-launch times are serialised and cold-cache; the roofline leg (dim 64, batch 4096: grids of 416 blocks of ids) "
            "runs first, then the metric leg (dim 128, batch 8192)\n")
    for k, v in agg.items():
        f.write("%-52s grid %-16s block %-12s n=%3d avg=%7.1f us min=%7.1f max=%7.1f\n" % (k[0][:52], k[1], k[2], len(v), sum(v) / len(v), min(v), max(v)))
raw = subprocess.run(["ncu", "-i", "gpurun_out/r2_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
want = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "smsp__inst_executed.sum": "warp_instructions",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak", "launch__waves_per_multiprocessor": "waves"}
stall = [i for i, h in enumerate(hdr) if "issue_stalled" in h and "not_issued" not in h and "per_issue" not in h]
full = []
for r in rr[2:]:
    d = {"kernel": re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])}
    for k, name in want.items():
        if k in hdr:
            d[name] = r[hdr.index(k)] + " " + units[hdr.index(k)]
    top = sorted(((float(r[i].replace(",", "")) if r[i] else 0.0, hdr[i].split("issue_stalled_")[-1]) for i in stall), reverse=True)[:4]
    d["top_stalls"] = [n for _, n in top]
    full.append(d)
json.dump({"command": "ncu --set full --clock-control none --import-source on (one launch per kernel) python bench.py --steps 20 "
                      "--warmup 10 --batch 4096 --dim 64 --no-cpu-baseline --no-graph --no-parity --no-model-leg --no-roofline-leg "
                      "--no-staleness", "kernels": full}, open("profiles/r2_ncu_full.json", "w"), indent=1)
print(open("profiles/r2_launch_list.txt").read())
for d in full:
    print(d)

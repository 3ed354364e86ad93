"""Summarise an `ncu --set full` report: per kernel duration, DRAM bytes (the roofline `traffic`), occupancy,
registers and the top stall reasons.  usage: summarize.py report.ncu-rep [out.json]"""
import csv
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__waves_per_multiprocessor": "waves",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__cycles_elapsed.max": "cycles_elapsed",
}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "ns": 1e-3, "usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3}


def main(rep, out=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for m, name in WANT.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                d[name] = v * UNIT.get(units[i], 1)
        if "dram_read" in d and "dram_write" in d:
            d["dram_bytes"] = d["dram_read"] + d["dram_write"]
        stalls = [(float(r[i] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                  for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
        d["top_stalls"] = [f"{n}:{v:.1f}" for v, n in sorted(stalls, reverse=True)[:4]]
        res.append(d)
    for d in res:
        print(json.dumps(d))
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)

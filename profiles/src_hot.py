"""Top stall lines of an ncu report (source page).  usage: src_hot.py report.ncu-rep [N]"""
import csv
import subprocess
import sys


def main(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    while rows and "Source" not in rows[0]:
        rows.pop(0)
    hdr = rows[0]
    ci = hdr.index("Source")
    samp = None
    for name in ("# Samples", "Warp Stall Sampling (All Samples)", "Warp Stall Sampling (All Cycles)"):
        if name in hdr:
            samp = hdr.index(name)
            break
    if samp is None:
        print(hdr)
        return
    data = []
    for r in rows[1:]:
        try:
            v = float(r[samp])
        except (ValueError, IndexError):
            v = 0
        data.append((v, r[ci][:120]))
    tot = sum(v for v, _ in data) or 1
    order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
    for i in order:
        v, sx = data[i]
        print(f"{i:5d} {v:8.0f} {100*v/tot:5.1f}%  {sx}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)

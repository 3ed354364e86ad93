"""Summarise an ncu `--metrics gpu__time_duration.sum` CSV: per-kernel durations of the last timed step."""
import csv
import re
import sys


def main(path, last=20):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tail = rows[-last:]
    tot = 0.0
    for x in tail:
        name = re.sub(r"\(.*", "", x["Kernel Name"])
        ns = float(x["Metric Value"])
        tot += ns
        print(f"{x['ID']:>6} {name[:64]:<64} {ns/1e3:9.2f} us  grid={x['Grid Size']} block={x['Block Size']}")
    print(f"sum of the listed launches: {tot/1e3:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20)

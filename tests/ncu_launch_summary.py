"""Per-kernel summary of an ncu launch list (measurement aid):
    ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file launches.csv <cmd>
    python tests/ncu_launch_summary.py launches.csv "<header line>" > profiles/<round>_launches_summary.txt
Times under ncu are cold-cache and serialised: compare SHARES with the bench's phase times, not absolutes."""
import csv
import sys
from collections import defaultdict


def main(path, header):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            rows.append((r["Kernel Name"], v))
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    total = sum(v for _, v in rows)
    print(header)
    print("%-86s %5s %10s %10s %6s" % ("kernel", "n", "mean us", "total us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-86s %5d %10.1f %10.1f %5.1f%%" % (k[:86], n, t / n, t, 100 * t / total))
    print("%-86s %5d %10s %10.1f" % ("total", len(rows), "", total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, count and mean / min / max us.
Consecutive launches are also grouped per process phase when --split N is given (every N launches of ufv kernels)."""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v * 1e6 if u == "s" else v
        rows.append((row["Kernel Name"], v, row.get("Grid Size", "")))
    return rows


def main():
    rows = load(sys.argv[1])
    agg = collections.OrderedDict()
    for name, v, grid in rows:
        key = (name.split("(")[0][-58:], grid)
        agg.setdefault(key, []).append(v)
    for (k, grid), v in agg.items():
        print(f"{k:58s} grid={grid:>14s} n={len(v):3d} mean={sum(v) / len(v):9.2f} us min={min(v):9.2f} max={max(v):9.2f}")


if __name__ == "__main__":
    main()

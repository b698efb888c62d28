"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
time and share.  Usage: python tools/ncu_summary.py launches.csv [top]"""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
        out.append((row["Kernel Name"], v))
    return out


def main():
    rows = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in rows)
    print(f"# {len(rows)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares, not absolutes)")
    print(f"# {'us':>10} {'share':>6} {'n':>5}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:12.1f} {100 * t / tot:5.1f}% {c:5d}  {k[:120]}")


if __name__ == "__main__":
    main()

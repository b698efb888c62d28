"""Top stall lines of one launch from `ncu -i rep --page source --csv [--launch-skip i --launch-count 1]`.
Usage: ncu -i rep.ncu-rep --page source --csv ... | python tools/ncu_hot_sass.py [top]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) >= len(hdr) - 2:
        data.append(r)
col = {h: i for i, h in enumerate(hdr)}
si = col["Warp Stall Sampling (All Samples)"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in data)
print(f"# {len(data)} SASS lines, {tot} samples")
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("# by reason: " + ", ".join(f"{k[6:]}={v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[si] or 0))[:top]:
    n = int(r[si] or 0)
    why = sorted(((int(r[col[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{100 * n / max(tot, 1):5.1f}%  {r[col['Source']].strip()[:90]:90s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")

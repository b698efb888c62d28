"""DRAM traffic per launch of one kernel family from an `ncu --set full` report -> profiles/dense_traffic.json, stamped
with the SHA-256 of the kernel's source so that bench.py refuses a capture taken from different code.
    ncu -i gpurun_out/r2_dense_full.ncu-rep --page raw --csv | python tools/ncu_traffic.py dense_tf32x3_kernel dense_tc.cu "<how>"
"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
family, source, how = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu --set full"
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, key):
    v = float(r[col[key]].replace(",", ""))
    u = units[col[key]].lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v * mult


tot, n = 0.0, 0
for r in rows[2:]:
    if family not in r[col["Kernel Name"]]:
        continue
    tot += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    n += 1
path = os.path.join(ROOT, "profiles", "dense_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[family] = {"source": source, "source_sha256": hashlib.sha256(open(os.path.join(ROOT, "dispu_b200", "csrc", source), "rb").read()).hexdigest(),
             "launches": n, "dram_bytes_read_plus_write_total": int(tot), "dram_bytes_per_launch": int(tot / max(n, 1)), "how": how}
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(d[family]))

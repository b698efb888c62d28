"""Per-kernel SASS instruction census of libdispu_b200.so (cuobjdump -sass): the Blackwell-native mnemonics the
profiling recipe names (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = TMA bulk
copy, STAS = st.async, SYNCS = mbarrier, UCGABAR = cluster barrier) next to FFMA2 (packed FP32), REDUX and total size.
Usage: python tools/sass_census.py [libdispu_b200.so] > profiles/r02_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dispu_b200", "libdispu_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
cols = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "LDGSTS", "SYNCS", "STAS", "UCGABAR", "FFMA2", "REDUX", "HMMA"]
kern = None
rows = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        rows[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if kern and m:
        op = m.group(1)
        rows[kern]["total"] += 1
        for c in cols:
            if op.startswith(c) or (c == "REDUX" and op.startswith("CREDUX")):
                rows[kern][c] += 1
print(f"# SASS census of {os.path.basename(so)} (sm_100a), one row per kernel: cuobjdump -sass | tools/sass_census.py")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier ops,")
print("# STAS = st.async (DSMEM), UCGABAR = cluster barrier, FFMA2 = packed fp32 FMA, REDUX = redux.sync, LDGSTS = cp.async (gather landing tiles),")
print("# HMMA = warp-level mma.sync: only the 16-sample weight-net contraction in the dense kernel's epilogue (m16n8k16 f16, scaled hi/lo split)\n")
print("| kernel | " + " | ".join(cols) + " | instr |")
print("|---|" + "---|" * (len(cols) + 1))
tot = collections.Counter()
for k, c in rows.items():
    name = demangle(k)
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("dispu::", "")
    print(f"| `{name[:70]}` | " + " | ".join(str(c[x]) if c[x] else "" for x in cols) + f" | {c['total']} |")
    tot.update(c)
print("| **all kernels** | " + " | ".join(str(tot[x]) for x in cols) + f" | {tot['total']} |")

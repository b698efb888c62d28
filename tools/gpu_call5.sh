#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
DISPU_STREAMS=0 DISPU_SPLIT_256=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $o/r2_bench_g.json 2> $o/r2_bench_g.err
python - <<P
import json
d=json.loads(open("$o/r2_bench_g.json").read().strip().splitlines()[-1])
print("g", round(d["ms_per_step"],4), round(d["value"]/1e6,3))
for k,v in d["roofline"]["all"].items():
    if "dense" in k: print(f"{k[:75]:75s} {v['ms_per_step']*1000:8.1f} x{v['launches_per_step']:.0f}")
P

#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
DISPU_STREAMS=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_h.json 2> $o/r2_bench_h.err
DISPU_STREAMS=0 DISPU_CHAIN=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_i.json 2> $o/r2_bench_i.err
for f in h i; do python - <<P
import json
try:
    d=json.loads(open("$o/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"])
    if "$f"=="h":
        for k,v in d["roofline"]["all"].items(): print(f"{k[:75]:75s} {v['ms_per_step']*1000:8.1f} x{v['launches_per_step']:.0f}")
except Exception as e:
    print("$f failed", e); print(open("$o/r2_bench_$f.err").read()[-1500:])
P
done

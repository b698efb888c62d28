#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "dense_conv or attention or linear_small or generator or smoke or mlp_chain" > $o/r2_t8.log 2>&1; echo "tests rc=$?" >> $o/r2_t8.log
tail -5 $o/r2_t8.log
DISPU_STREAMS=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_j.json 2> $o/r2_bench_j.err
DISPU_STREAMS=0 DISPU_EDGE_SIMT=1 DISPU_ATTENTION_PACK=generic timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_k.json 2> $o/r2_bench_k.err
for f in j k; do python - <<P
import json
try:
    d=json.loads(open("$o/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"])
    if "$f"=="j":
        for k,v in d["roofline"]["all"].items(): print(f"{k[:75]:75s} {v['ms_per_step']*1000:8.1f} x{v['launches_per_step']:.0f}")
except Exception as e:
    print("$f failed", e); print(open("$o/r2_bench_$f.err").read()[-1500:])
P
done

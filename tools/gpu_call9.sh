#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $o/r2_t9.log 2>&1; echo "tests rc=$?" >> $o/r2_t9.log
tail -4 $o/r2_t9.log
DISPU_STREAMS=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_l.json 2> $o/r2_bench_l.err
DISPU_STREAMS=0 DISPU_NO_PDL=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_m.json 2> $o/r2_bench_m.err
for f in l m; do python - <<P
import json
try:
    d=json.loads(open("$o/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"], d["loss"])
except Exception as e:
    print("$f failed", e); print(open("$o/r2_bench_$f.err").read()[-1500:])
P
done

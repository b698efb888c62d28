#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "dense or generator" > $o/r2_t3.log 2>&1; echo "tests rc=$?" >> $o/r2_t3.log
timeout 120 python tools/knn_diag.py > $o/r2_knn_diag_tma.log 2>&1
DISPU_NO_TMA=1 timeout 120 python tools/knn_diag.py > $o/r2_knn_diag_notma.log 2>&1
DISPU_STREAMS=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_d.json 2> $o/r2_bench_d.err
DISPU_STREAMS=0 DISPU_DENSE_NO_PAIR=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_e.json 2> $o/r2_bench_e.err
tail -3 $o/r2_t3.log; cat $o/r2_knn_diag_tma.log $o/r2_knn_diag_notma.log | tail -12
for f in d e; do python - <<P
import json
try:
    d=json.loads(open("$o/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"])
except Exception as e:
    print("$f failed", e)
P
done

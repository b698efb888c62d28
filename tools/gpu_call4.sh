#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $o/r2_t4.log 2>&1; echo "tests rc=$?" >> $o/r2_t4.log
DISPU_STREAMS=0 timeout 300 python bench.py --steps 30 --warmup 5 > $o/r2_bench_f.json 2> $o/r2_bench_f.err
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 > $o/r2_bench_c3.json 2> $o/r2_bench_c3.err
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 > $o/r2_bench_c4.json 2> $o/r2_bench_c4.err
timeout 600 python bench.py --config c5 --steps 3 --warmup 3 > $o/r2_bench_c5.json 2> $o/r2_bench_c5.err
timeout 120 python tools/knn_diag.py > $o/r2_knn_diag2.log 2>&1
tail -3 $o/r2_t4.log; tail -4 $o/r2_knn_diag2.log
for f in f c3 c4 c5; do python - <<P
import json
try:
    d=json.loads(open("$o/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"], d["roofline"]["kernel"][:40], round(d["roofline"]["frac"],4))
except Exception as e:
    print("$f failed", e); print(open("$o/r2_bench_$f.err").read()[-1500:])
P
done

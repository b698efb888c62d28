#!/bin/bash
# development: one GPU-box session = tests + A/B benches (outputs under gpurun_out/)
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $o/r2_t_all2.log 2>&1; echo "tests rc=$?" >> $o/r2_t_all2.log
timeout 120 python tools/bench_attention.py > $o/r2_att_bench2.log 2>&1
timeout 200 python tests/golden/make_golden_real.py gpu > $o/r2_golden_real.log 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 > $o/r2_bench_a.json 2> $o/r2_bench_a.err
DISPU_STREAMS=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_b.json 2> $o/r2_bench_b.err
DISPU_STREAMS=0 DISPU_ATTENTION=two_products timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/r2_bench_c.json 2> $o/r2_bench_c.err
timeout 300 python tools/microbench.py --quick > $o/r2_micro_tma.log 2>&1
DISPU_NO_TMA=1 timeout 300 python tools/microbench.py --quick > $o/r2_micro_notma.log 2>&1
tail -3 $o/r2_t_all2.log
for f in a b c; do python - <<P
import json
try:
    d=json.loads(open("$o/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"])
except Exception as e:
    print("$f failed", e)
P
done

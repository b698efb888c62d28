#!/bin/bash
# Evidence session on ONE B200: full GPU test suite, smoke, bench lines of every config + the reference arm, ncu launch list of
# the c2 step, ncu --set full of the dense kernels and of the other kernels of the step (reports under gpurun_out/, read on
# the CPU box and summarised under profiles/).
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $o/ev_gputests.log 2>&1; echo "tests rc=$?" >> $o/ev_gputests.log; tail -3 $o/ev_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/ev_smoke.log 2>&1; tail -1 $o/ev_smoke.log
timeout 300 python bench.py --steps 30 --warmup 5 > $o/ev_c2_n1.json 2> $o/ev_c2_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $o/ev_ref.json 2> $o/ev_ref.err
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 > $o/ev_c3_n1.json 2> $o/ev_c3_n1.err
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 > $o/ev_c4_n1.json 2> $o/ev_c4_n1.err
timeout 600 python bench.py --config c5 --steps 3 --warmup 3 > $o/ev_c5_n1.json 2> $o/ev_c5_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/ev_launches.csv python tools/run_step_once.py 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tf32x3_kernel -s 11 -c 11 -f -o $o/ev_dense_full python tools/run_step_once.py 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:mlp_chain_kernel|attention_tf32x3_kernel|edge_mlp_tc_kernel|feature_knn_kernel|knn_select|nn_min_reduce_kernel|attention_pack_kernel|group_max_kernel" -s 14 -c 14 -f -o $o/ev_others_full python tools/run_step_once.py 2 > /dev/null 2>&1
ls -la $o/ev_*.ncu-rep | tail -3
for f in c2 c3 c4 c5; do python - <<P
import json
try:
    d=json.loads(open("$o/ev_${f}_n1.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"], d["roofline"]["kernel"][:30], round(d["roofline"]["frac"],4), d["clocks"])
except Exception as e:
    print("$f failed", e)
P
done
tail -c 400 $o/ev_ref.json

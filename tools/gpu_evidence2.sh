#!/bin/bash
# Refresh of the c2 evidence after the last kernel changes: GPU tests, smoke, c2 / c4 bench lines, launch list, dense + others ncu
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $o/ev_gputests.log 2>&1; echo "tests rc=$?" >> $o/ev_gputests.log; tail -3 $o/ev_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $o/ev_smoke.log 2>&1; tail -1 $o/ev_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/ev_launches.csv python tools/run_step_once.py 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tf32x3_kernel -s 11 -c 11 -f -o $o/ev_dense_full python tools/run_step_once.py 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:mlp_chain_kernel|attention_tf32x3_kernel|edge_mlp_tc_kernel|feature_knn_kernel|knn_select|nn_min_reduce_kernel|attention_pack_kernel|group_max_kernel" -s 14 -c 14 -f -o $o/ev_others_full python tools/run_step_once.py 2 > /dev/null 2>&1
ls -la $o/ev_*.ncu-rep | tail -3

#!/bin/bash
# Refresh of the dense-kernel evidence after a change to dense_tc.cu: dense tests, ncu --set full of the dense launches of one
# c2 step (-> profiles/r02_dense_tc_ncu.md and the traffic stamp, tools/ncu_traffic.py), one c2 bench line
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "dense or generator or weightnet" > $o/ev_densetests.log 2>&1; echo "tests rc=$?" >> $o/ev_densetests.log; tail -3 $o/ev_densetests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tf32x3_kernel -s 11 -c 11 -f -o $o/ev_dense_full python tools/run_step_once.py 2 > /dev/null 2>&1
ls -la $o/ev_dense_full.ncu-rep

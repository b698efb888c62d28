#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none -k regex:dense_tf32x3 -s 2 -c 1 -f -o $o/s2_dense_lean python tools/run_dense_once.py 32768x2048x256 > $o/s2_ncu.log 2>&1
tail -3 $o/s2_ncu.log; ls -la $o/*.ncu-rep

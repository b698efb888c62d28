#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none -k regex:dense_tf32x3 -s 4 -c 1 -f -o $o/s2_conv1 python tools/conv1_dbg.py > $o/s2_ncu.log 2>&1
tail -2 $o/s2_ncu.log; ls -la $o/*.ncu-rep

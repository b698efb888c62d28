#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 60 python tools/conv1_dbg.py 2>&1 | tail -1
DISPU_WN_EPI=ffma timeout 60 python tools/conv1_dbg.py 2>&1 | tail -1
DISPU_WN_EPI=ffma timeout 100 python -m pytest tests -m gpu -q -x -k "weightnet or generator" 2>&1 | tail -2
DISPU_WN_EPI=ffma timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['loss'])"

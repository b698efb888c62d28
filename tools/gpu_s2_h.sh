#!/bin/bash
cd "$(dirname "$0")/.."
DISPU_SPLIT_256=1 timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['loss'])
for k,v in d['roofline']['all'].items():
    if 'dense_tf32x3' in k: print(k, round(v['ms_per_launch']*1e3,1))"

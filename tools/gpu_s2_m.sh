#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "feature or dense_conv or generator or edge" 2>&1 | tail -3
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/s2c_c2.json 2> $o/s2c_c2.err
python - <<P
import json
d=json.loads(open("$o/s2c_c2.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["loss"])
for k,v in d["roofline"]["all"].items():
    print(k, round(v["ms_per_launch"]*1e3,1))
P

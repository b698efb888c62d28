#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x -k "knn or generator or real_cloud or tma" 2>&1 | tail -3
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/s2c_c2.json 2> $o/s2c_c2.err
python - <<P
import json
d=json.loads(open("$o/s2c_c2.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"], d["roofline"]["frac"], d["roofline"].get("traffic"), d["loss"])
for k,v in d["roofline"]["all"].items():
    print(k, round(v["ms_per_launch"]*1e3,1))
P

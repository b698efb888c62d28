#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x -k "dense or generator or attention or chain or linear or weight or feature or edge or knn" > $o/s2c_tests.log 2>&1; echo "tests rc=$?" >> $o/s2c_tests.log
tail -4 $o/s2c_tests.log
timeout 60 python tools/conv1_dbg.py 2>&1 | tail -1
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $o/s2c_c2.json 2> $o/s2c_c2.err
python - <<P
import json
d=json.loads(open("$o/s2c_c2.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"], d["roofline"]["frac"], d["loss"])
for k,v in d["roofline"]["all"].items():
    print(k, round(v["ms_per_launch"]*1e3,1))
P

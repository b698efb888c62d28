#!/bin/bash
# session-2 call A: validate HEAD (tests + c2 bench with per-kernel in-step times)
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $o/s2a_tests.log 2>&1; echo "tests rc=$?" >> $o/s2a_tests.log
tail -4 $o/s2a_tests.log
timeout 300 python bench.py --steps 30 --warmup 5 > $o/s2a_c2.json 2> $o/s2a_c2.err
python - <<P
import json
d=json.loads(open("$o/s2a_c2.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"], d["roofline"]["frac"], d["roofline"].get("traffic"), d["clocks"])
for k,v in sorted(d["roofline"]["all"].items(), key=lambda kv:-kv[1].get("us_per_step",0) if isinstance(kv[1],dict) else 0):
    print(k, v)
P

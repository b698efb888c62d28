#!/bin/bash
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 > $o/ev_c2_n1.json 2> $o/ev_c2_n1.err
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 > $o/ev_c4_n1.json 2> $o/ev_c4_n1.err
for f in c2 c4; do python - <<P
import json
d=json.loads(open("$o/ev_${f}_n1.json").read().strip().splitlines()[-1])
print("$f", round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"], round(d["roofline"]["frac"],4), d["roofline"].get("traffic"), d["clocks"])
P
done

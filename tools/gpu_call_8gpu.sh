#!/bin/bash
# 8-GPU session: c2 / c4 / c5 through the driver's own launch line (torch.distributed.run, one rank per GPU)
# usage: gpu_call_8gpu.sh [configs...]   (default: c2 c2nccl c4 c5)
cd "$(dirname "$0")/.."
o=gpurun_out
run() { # name config steps warmup extra
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --config $2 --steps $3 --warmup $4 --no-cpu-baseline $5 > $o/ev_n8_$1.json 2> $o/ev_n8_$1.err
  echo "$1 rc=$?"
}
sel="${@:-c2 c2nccl c4 c5}"
for s in $sel; do
  case $s in
    c2) run c2 c2 30 5 ;;
    c2nccl) run c2nccl c2 30 5 "--collective nccl" ;;
    c4) run c4 c4 10 3 ;;
    c5) run c5 c5 3 3 ;;
  esac
done
python - <<P
import json
for f in "$sel".split():
    try:
        d=json.loads(open("$o/ev_n8_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["config"].get("sharding"))
    except Exception as e:
        print(f, "failed", e)
P

#!/bin/bash
# 8-GPU session: c2 / c4 / c5 through the driver's own launch line (torch.distributed.run, one rank per GPU)
cd "$(dirname "$0")/.."
o=gpurun_out
run() { # name config steps warmup
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --config $2 --steps $3 --warmup $4 --no-cpu-baseline > $o/r2_n8_$1.json 2> $o/r2_n8_$1.err
  echo "$1 rc=$?"; tail -c 600 $o/r2_n8_$1.json | cut -c1-400; grep -c "NCCL INFO.*Init COMPLETE\|ncclCommInitRank.*nranks 8" $o/r2_n8_$1.err
}
run c2 c2 30 5
run c4 c4 10 3
run c5 c5 3 3
DISPU_STREAMS=0 timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > $o/r2_n1_c4.json 2> $o/r2_n1_c4.err
python - <<P
import json
for f in ["n8_c2","n8_c4","n8_c5","n1_c4"]:
    try:
        d=json.loads(open("$o/r2_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["config"].get("sharding"))
    except Exception as e:
        print(f, "failed", e)
P

#!/bin/bash
# N-GPU session for the peer-memory all-reduce: its tests, then c2 with the library's kernel vs ncclAllReduce
cd "$(dirname "$0")/.."
o=gpurun_out; N=${1:-2}
timeout 600 python -m pytest tests/test_peer_reduce.py -q -x 2>&1 | tail -5
for coll in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --config c2 --steps 50 --warmup 5 --no-cpu-baseline --collective $coll > $o/peer_n${N}_$coll.json 2> $o/peer_n${N}_$coll.err
  echo "$coll rc=$?"; tail -3 $o/peer_n${N}_$coll.err | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_collective.py 2>&1 | grep world
timeout 300 python bench.py --config c2 --steps 50 --warmup 5 --no-cpu-baseline > $o/peer_n1.json 2> $o/peer_n1.err
python - <<P
import json
for f in ["n${N}_peer","n${N}_nccl","n1"]:
    try:
        d=json.loads(open("$o/peer_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["ms_per_step"],4), round(d["value"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3), d["gpu_launches_per_step"], d["config"].get("sharding"))
    except Exception as e:
        print(f, "failed", e)
P

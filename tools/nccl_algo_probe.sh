#!/bin/bash
# 8-GPU probe: which NCCL algorithm carries the policy all-gather, and does forcing NVLS / a protocol change the exchange-bound tick?
OUT=gpurun_out; mkdir -p $OUT
run() { name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 8 --steps 8 --warmup 3 --no-extra > $OUT/nccl_$name.json 2> $OUT/nccl_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/nccl_$name.json").read().strip().splitlines()[-1]); print("$name", round(d["value"]), round(d["ms_per_step"], 2))
except Exception as e:
    print("$name failed", e)
PY
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,COLL
grep -iE "allgather|nvls|algo" $OUT/nccl_default.err | grep -v "^$" | sort | uniq -c | sort -rn | head -12
run nvls NCCL_ALGO=NVLS
run ring NCCL_ALGO=Ring

#!/bin/bash
# Multi-GPU pass on ONE box with N GPUs: policy-exchange variants of bench.py (weak scaling, 4096 instances per GPU).
# usage (under gpurun --gpus N): bash tools/scale_session.sh <tag> <N> [variants...]
TAG=${1:-r02s}; N=${2:-2}; shift 2
OUT=gpurun_out; mkdir -p $OUT
VARIANTS=("$@")
if [ ${#VARIANTS[@]} -eq 0 ]; then VARIANTS=("native16" "nativece" "torch0" "none"); fi
run() {  # name, extra args
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 10 --warmup 3 "$@" > $OUT/${TAG}_n${N}_${name}.json 2> $OUT/${TAG}_n${N}_${name}.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_n${N}_${name}.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "phases", {k: round(v, 2) for k, v in d["phase_ms"].items()}, d.get("policy_exchange"))
    for k, v in (d.get("other_configs") or {}).items():
        print("   ", k, round(v.get("value", 0)) if isinstance(v, dict) else v, v.get("ms_per_step") if isinstance(v, dict) else "")
except Exception as e:
    print("no result:", e); print(open("$OUT/${TAG}_n${N}_${name}.err").read()[-1500:])
PY
}
for v in "${VARIANTS[@]}"; do
  case $v in
    native16) run $v --gather-impl native --gather-ctas 16 --no-extra ;;
    native8) run $v --gather-impl native --gather-ctas 8 --no-extra ;;
    native32) run $v --gather-impl native --gather-ctas 32 --no-extra ;;
    native0) run $v --gather-impl native --gather-ctas 0 --no-extra ;;
    nativece) run $v --gather-impl native --gather-ctas 0 --gather-ce 1 --no-extra ;;
    sym8) run $v --gather-impl native --gather-ctas 8 --gather-ce 2 --no-extra ;;
    sym16) run $v --gather-impl native --gather-ctas 16 --gather-ce 2 --no-extra ;;
    sym32) run $v --gather-impl native --gather-ctas 32 --gather-ce 2 --no-extra ;;
    sym0) run $v --gather-impl native --gather-ctas 0 --gather-ce 2 --no-extra ;;
    torch0) run $v --gather-impl torch --gather-ctas 0 --no-extra ;;
    torch16) run $v --gather-impl torch --gather-ctas 16 --no-extra ;;
    none) run $v --gather none --no-extra ;;
    full) run $v ;;
  esac
done
ls -la $OUT | tail -12

#!/bin/bash
# bench every variants/libbmpc_*.so (compile-time tuning variants) on the GPU box; prints value + phase times per variant
mkdir -p gpurun_out
cp bipedal_control_b200/libbmpc.so /tmp/libbmpc_keep.so
for f in variants/libbmpc_*.so; do
  cp $f bipedal_control_b200/libbmpc.so
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline $BENCH_OPTS > gpurun_out/var.json 2> gpurun_out/var.err
  python - "$f" <<'PY'
import json, sys
try:
    j = json.load(open("gpurun_out/var.json"))
    print(sys.argv[1], "value %.0f ms %.3f" % (j["value"], j["ms_per_step"]), {k: round(v, 3) for k, v in j["phase_ms"].items()}, "status", j["status_nonzero"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/var.err").read()[-400:])
PY
done
cp /tmp/libbmpc_keep.so bipedal_control_b200/libbmpc.so

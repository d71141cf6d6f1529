#!/bin/bash
# quick GPU pass: parity tests + bench line (no ncu).  usage: bash tools/gpu_quick.sh <tag> [pytest -k expression]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$2" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > $OUT/${TAG}_pytest_gpu.log 2>&1
else timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; fi
echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
tail -15 $OUT/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
j=json.load(open("$OUT/${TAG}_bench.json"))
print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], "status_nonzero", j["status_nonzero"])
print(j["phase_ms"])
PY
tail -3 $OUT/${TAG}_bench.err

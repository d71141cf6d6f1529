#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu --set full of the heavy kernels (CSV pages exported on the box).
# usage (under gpurun): bash tools/gpu_profile.sh <tag> [skip-tests]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  tail -3 $OUT/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 3000 $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
KREG='regex:k_lq_pack|k_riccati_warp|k_project|k_policy_expand|k_linesearch_eval2|k_forward'
timeout 900 ncu --set full --clock-control none --import-source on -k "$KREG" --launch-skip 12 --launch-count 6 -o $OUT/${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_full.ncu-rep --page details --csv > $OUT/${TAG}_full_details.csv 2>/dev/null
ncu -i $OUT/${TAG}_full.ncu-rep --page source --csv > /tmp/${TAG}_full_source.csv 2>/dev/null
# the source page prints every kernel twice: keep the first copy of each, drop the .ncu-rep (gpurun_out is limited to 64 MiB)
python - <<PY
import re
src = open("/tmp/${TAG}_full_source.csv").read().split('"Kernel Name",')
seen, out = set(), []
for s in src[1:]:
    k = s.split("\n", 1)[0]
    if k in seen:
        continue
    seen.add(k)
    out.append('"Kernel Name",' + s)
open("$OUT/${TAG}_full_source.csv", "w").write("".join(out))
PY
rm -f $OUT/${TAG}_full.ncu-rep
ls -la $OUT

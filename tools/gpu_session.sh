#!/bin/bash
# One GPU-box pass (1 GPU): parity tests, smoke, bench line, ncu launch list, ncu --set full of the heavy kernels (CSV pages exported on the box).
# usage (under gpurun): bash tools/gpu_session.sh <tag> [notest] [noprof]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [[ " $* " != *" notest "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_gpu.log
  tail -15 $OUT/${TAG}_pytest_gpu.log
  if ! grep -q "pytest exit 0" $OUT/${TAG}_pytest_gpu.log; then
    timeout 300 python tools/dev_compare.py 2 3 > $OUT/${TAG}_dev_compare.log 2>&1; tail -60 $OUT/${TAG}_dev_compare.log
  fi
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 4000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
if [[ " $* " != *" noprof "* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $OUT/${TAG}_ncu_launch.log 2>&1
  KREG='regex:k_lq_pack|k_riccati_warp|k_project|k_policy_expand|k_linesearch|k_forward'
  timeout 900 ncu --set full --clock-control none --import-source on -k "$KREG" --launch-skip 12 --launch-count 6 -o $OUT/${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $OUT/${TAG}_ncu_full.log 2>&1
  ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_full.ncu-rep --page details --csv > $OUT/${TAG}_full_details.csv 2>/dev/null
  ncu -i $OUT/${TAG}_full.ncu-rep --page source --csv > /tmp/${TAG}_full_source.csv 2>/dev/null
  # the source page prints every kernel twice: keep the first copy of each, drop the .ncu-rep (gpurun_out is limited to 64 MiB)
  python - <<PY
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
fi
ls -la $OUT

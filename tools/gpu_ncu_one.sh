#!/bin/bash
# ncu --set full of one kernel regex (warm tick), CSV pages exported.  usage: bash tools/gpu_ncu_one.sh <tag> <kernel-regex> [launch-skip]
TAG=$1; KREG=$2; SKIP=${3:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KREG" --launch-skip $SKIP --launch-count 1 -o $OUT/${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $OUT/${TAG}_full.ncu-rep --page details --csv > $OUT/${TAG}_full_details.csv 2>/dev/null
ncu -i $OUT/${TAG}_full.ncu-rep --page source --csv > $OUT/${TAG}_full_source.csv 2>/dev/null
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
rm -f $OUT/${TAG}_full.ncu-rep
ls -la $OUT | grep $TAG

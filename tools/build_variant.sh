#!/bin/bash
# build a compile-time tuning variant of libbmpc.so into variants/ (benched on the GPU box by tools/gpu_variants.sh)
# usage: tools/build_variant.sh <name> [-DKNOB=value ...]
NAME=$1; shift
mkdir -p variants
cd bipedal_control_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -shared \
  -o ../../variants/libbmpc_${NAME}.so bmpc_api.cu bmpc_model.cpp bmpc_ingest.cpp -lcudart -ldl 2> ../../variants/${NAME}.log || { tail -20 ../../variants/${NAME}.log; exit 1; }

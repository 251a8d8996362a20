#!/bin/bash
# A/B of the GEMM timings on one box: tools/gpu_ab_gemm.sh <other.so> [grep pattern] [rounds]; whole processes alternate
other=$1; pat=${2:-.}; rounds=${3:-2}
for rep in $(seq $rounds); do
  for lib in $other neural_vit_b200/libtvit_b200.so; do
    echo "== $lib"
    TVIT_LIB_PATH=$PWD/$lib timeout 300 python tools/bench_kernels.py --only gemm --reps 10 2>&1 | grep -E "$pat"
  done
done

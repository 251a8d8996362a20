#!/bin/bash
# A/B of the attention kernels on one box: tools/gpu_ab_attn.sh <other.so> [grep pattern] [rounds]
# (<other.so> from tools/build_ab.sh or a copy of an earlier build).  Whole processes alternate -- other, in-tree, other,
# in-tree -- because under the power cap the clock sags within a process and biases whatever is timed last.
other=$1; pat=${2:-attn}; rounds=${3:-2}
for rep in $(seq $rounds); do
  for lib in $other neural_vit_b200/libtvit_b200.so; do
    echo "== $lib"
    TVIT_LIB_PATH=$PWD/$lib timeout 300 python tools/bench_kernels.py --only attn --reps 10 2>&1 | grep -E "$pat"
  done
done

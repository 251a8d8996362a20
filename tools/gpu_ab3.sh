#!/bin/bash
# A/B of the attention backward timings: $1 = other library; interleaved with the in-tree one
for rep in 1 2; do
  for lib in $1 neural_vit_b200/libtvit_b200.so; do
    echo "== $lib"
    TVIT_LIB_PATH=$PWD/$lib timeout 300 python tools/bench_kernels.py --only attn --reps 10 2>&1 | grep -E "attn bwd drop"
  done
done

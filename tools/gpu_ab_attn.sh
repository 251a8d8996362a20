#!/bin/bash
# A/B of the attention kernels on one box: ab/libtvit_prev.so (previous build) vs the in-tree library, interleaved
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention or attn" 2>&1 | tail -n 3
for rep in 1 2; do
  for lib in ab/libtvit_prev.so neural_vit_b200/libtvit_b200.so; do
    echo "== $lib"
    TVIT_LIB_PATH=$PWD/$lib timeout 300 python tools/bench_kernels.py --only attn --reps 10 2>&1 | grep -E "attn" 
  done
done

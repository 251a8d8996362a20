#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention or attn" 2>&1 | tail -n 15
timeout 300 python tools/bench_kernels.py --only attn --reps 10 2>&1 | grep -E "attn"

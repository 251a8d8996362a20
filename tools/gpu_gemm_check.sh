#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "gemm or replay or layernorm or branch" 2>&1 | tail -n 4
timeout 900 python bench.py --steps 3 --warmup 2 --dropout 0 --no-cpu-baseline --breakdown > gpurun_out/bench_nodrop.log 2> gpurun_out/bench_nodrop.err; tail -n 2 gpurun_out/bench_nodrop.err
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --breakdown > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -n 2 gpurun_out/bench.err

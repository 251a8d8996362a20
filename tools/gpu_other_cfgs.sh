#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --layers 12 > gpurun_out/bench_c3_depth12.log 2>gpurun_out/c3.err; echo "c3 $?"
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --layers 24 --embed-dim 768 --heads 12 --batch 64 > gpurun_out/bench_c5_large.log 2>gpurun_out/c5.err; echo "c5 $?"
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --trials 32 --freq 128 --time 512 --batch 8 > gpurun_out/bench_c4_longseq.log 2>gpurun_out/c4.err; echo "c4 $?"
tail -n 2 gpurun_out/c3.err gpurun_out/c5.err gpurun_out/c4.err

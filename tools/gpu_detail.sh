#!/bin/bash
mkdir -p gpurun_out
TVIT_BENCH_DETAIL=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --breakdown > gpurun_out/bench_detail.log 2> gpurun_out/bench_detail.err
grep "per-op" gpurun_out/bench_detail.err | tr ',' '\n' | sed 's/\[//; s/\]//' | paste - - | sort | head -40

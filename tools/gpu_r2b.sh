#!/bin/bash
# full GPU suite + smoke + the bench step (with per-op breakdown) on the current tree
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --breakdown > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
cut -c1-1500 gpurun_out/bench.log; grep -E "attn|gemm|ln_|total|colsum" gpurun_out/bench.err | tail -n 30

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 2 --breakdown > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
timeout 600 python bench.py --steps 3 --warmup 2 --dropout 0 --no-cpu-baseline --breakdown > gpurun_out/bench_nodrop.log 2> gpurun_out/bench_nodrop.err; echo "bench nodrop exit $?"
cat gpurun_out/bench_nodrop.log; tail -n 3 gpurun_out/bench_nodrop.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "bench reference exit $?"
cut -c1-400 gpurun_out/bench_reference.log

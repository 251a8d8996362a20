#!/bin/bash
# usage: gpu_multi3.sh N config...  -- the named configs on N GPUs of one box (driver launch line)
N=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2b_bench_${c}_dp$N.json 2> gpurun_out/r2b_bench_${c}_dp$N.err; echo "bench $c dp$N exit $?"
  tail -n 1 gpurun_out/r2b_bench_${c}_dp$N.json | cut -c1-200
done

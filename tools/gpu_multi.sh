#!/bin/bash
# usage: gpu_multi.sh N
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/bench_dp$N.log 2> gpurun_out/bench_dp$N.err; echo "bench dp$N exit $?"
tail -n 3 gpurun_out/bench_dp$N.err; cat gpurun_out/bench_dp$N.log

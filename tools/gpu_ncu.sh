#!/bin/bash
mkdir -p gpurun_out
PROF_B=64 python tools/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
PROF_B=64 ncu --set full --clock-control none --import-source on -k regex:"tc_|ln_|colsum|branch" -o gpurun_out/prof_r1 -f python tools/prof_kernels.py > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -n 5 gpurun_out/ncu.log; ls -la gpurun_out/*.ncu-rep

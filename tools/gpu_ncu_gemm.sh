#!/bin/bash
mkdir -p gpurun_out
PROF_B=256 python tools/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
PROF_B=256 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm" -o gpurun_out/prof_gemm -f python tools/prof_kernels.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_gemm.log

#!/bin/bash
mkdir -p gpurun_out
PROF_B=256 PROF_DROPOUT=1 python tools/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 && \
PROF_B=256 PROF_DROPOUT=1 ncu --set full --clock-control none --import-source on -k tc_gemm_kernel -s 2 -c 3 -o gpurun_out/prof_gemm -f python tools/prof_kernels.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_gemm.log

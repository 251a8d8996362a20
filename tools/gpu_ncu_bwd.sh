#!/bin/bash
mkdir -p gpurun_out
export PROF_B=256 PROF_DROPOUT=1
python tools/prof_attn.py > gpurun_out/prof_attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k tc_attn_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_bwd -f python tools/prof_attn.py > gpurun_out/ncu_bwd.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_bwd.log

#!/bin/bash
# ncu --set full on the attention kernels at the bench shape, dropout on (after the same command ran plain and exited 0)
mkdir -p gpurun_out
export PROF_B=${PROF_B:-256} PROF_DROPOUT=${PROF_DROPOUT:-1}
python tools/prof_attn.py > gpurun_out/prof_attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tc_attn" -s 2 -c 2 -o gpurun_out/prof_attn -f python tools/prof_attn.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_attn.log

#!/bin/bash
# round-2 (second session) evidence: full GPU suite, smoke, ncu --set full of the attention kernels, launch list of the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
export PROF_B=256 PROF_DROPOUT=1
python tools/prof_attn.py > gpurun_out/prof_attn_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_attn" -s 2 -c 2 -o gpurun_out/prof_attn -f python tools/prof_attn.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_attn.log
ncu -i gpurun_out/prof_attn.ncu-rep --page raw --csv > gpurun_out/r2b_prof_attn_raw.csv 2>/dev/null; ls -la gpurun_out/prof_attn.ncu-rep gpurun_out/r2b_prof_attn_raw.csv
rm -f gpurun_out/prof_attn.ncu-rep
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/b_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 700 --csv --log-file gpurun_out/r2b_ncu_launchlist.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launchlist exit $?"; wc -l gpurun_out/r2b_ncu_launchlist.csv

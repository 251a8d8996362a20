#!/bin/bash
# final evidence of the round: full GPU suite, smoke, ncu --set full of the attention kernels, driver-flag bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
export PROF_B=256 PROF_DROPOUT=1
python tools/prof_attn.py > gpurun_out/prof_attn_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_attn" -s 2 -c 2 -o gpurun_out/prof_attn -f python tools/prof_attn.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/prof_attn.ncu-rep --page raw --csv > gpurun_out/r2b_prof_attn_raw.csv 2>/dev/null
rm -f gpurun_out/prof_attn.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_c2_final.json 2> gpurun_out/r2b_bench_c2_final.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_reference_arm.json 2> gpurun_out/r2b_bench_reference_arm.err; echo "reference arm exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_bench_c2_final.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms", d["clocks"], "e2e", round(d["e2e"]["value"], 1), "roofline", d["roofline"])
r = json.loads(open("gpurun_out/r2b_bench_reference_arm.json").read().strip().splitlines()[-1])
print("reference arm", r.get("value"), r.get("unit"), r.get("cpu_baseline", {}).get("cores"))
PY

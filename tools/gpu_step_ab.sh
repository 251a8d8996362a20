#!/bin/bash
# whole-step A/B on one box: env settings given as arguments ("VAR=a VAR=b ..."), alternating, two rounds
mkdir -p gpurun_out
for rep in 1 2; do
  i=0
  for setting in "$@"; do
    i=$((i+1))
    env $setting timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --breakdown > gpurun_out/bench_ab$i.log 2> gpurun_out/bench_ab$i.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/bench_ab$i.log").read().strip().splitlines()[-1])
print("$setting", round(d["value"], 1), "samples/s", round(d["ms_per_step"], 2), "ms", d["clocks"]["sm_mhz"], "MHz", "attn fwd/bwd", d["breakdown_ms"]["attn_fwd"], d["breakdown_ms"]["attn_bwd"], "roofline frac", round(d["roofline"]["frac"], 4))
PY
  done
done

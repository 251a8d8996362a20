#!/bin/bash
timeout 300 python tools/diag_gpu.py 2>&1 | grep -E "TFLOP|calib" | tail -n 8
timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --breakdown > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("dropout", round(d["value"], 1), round(d["ms_per_step"], 2), round(d.get("model_frac_of_peak"), 4), "e2e", round(d["e2e"]["value"], 1), d["clocks"])
PY
grep -E "per-op" gpurun_out/bench.err | tail -n 2

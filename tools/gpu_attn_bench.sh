#!/bin/bash
# attention parity + isolated timings, attention op tests, then the bench step with and without dropout
mkdir -p gpurun_out
timeout 600 python tools/diag_attn.py --bwd 2>&1 | grep -v Warn | tail -n 9
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -n 4
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --breakdown > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench.log",):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d.get("model_frac_of_peak"), d["roofline"]["ms_per_launch"])
PY
grep -E "attn|gemm|total" gpurun_out/bench.err | tail -n 16
timeout 600 python bench.py --steps 3 --warmup 2 --dropout 0 --no-cpu-baseline > gpurun_out/bench_nodrop.log 2> gpurun_out/bench_nodrop.err; echo "bench nodrop exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_nodrop.log").read().strip().splitlines()[-1])
print("nodrop", d["value"], d["ms_per_step"], d.get("model_frac_of_peak"), d["roofline"]["ms_per_launch"])
PY

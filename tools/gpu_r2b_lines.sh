#!/bin/bash
# second-session bench lines, one GPU: C2 (driver flags, full line with both baselines), C2 without dropout, C3, C4, C5
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "c2 exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 --dropout 0 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2b_bench_c2_nodrop.json 2>/dev/null; echo "c2 nodrop exit $?"
for c in c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2b_bench_$c.json 2> gpurun_out/r2b_bench_$c.err; echo "$c exit $?"
done
python - <<'PY'
import json
for n in ("c2", "c2_nodrop", "c3", "c4", "c5"):
    try:
        d = json.loads(open(f"gpurun_out/r2b_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms", "frac", round(d.get("model_frac_of_peak", 0), 4), d["clocks"]["sm_mhz"], "MHz", "e2e", round(d["e2e"]["value"], 1), "attn-bwd frac", round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(n, "FAILED", e)
PY

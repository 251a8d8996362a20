#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_mode.py -q -m gpu -x -k "attention or attn or train_mode" 2>&1 | tail -n 3
timeout 300 python tools/bench_kernels.py --only attn --reps 10 2>&1 | grep -E "attn (fwd|bwd) drop"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --breakdown 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['breakdown_ms']['attn_fwd'], d['breakdown_ms']['attn_bwd'])"

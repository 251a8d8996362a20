#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_attn.py --bwd > gpurun_out/diag_attn_bwd.log 2>&1; echo "diag_attn exit $?"
tail -n 40 gpurun_out/diag_attn_bwd.log

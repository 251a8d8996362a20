#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_attn.py --bwd 2>&1 | grep -v Warn | tail -n 9

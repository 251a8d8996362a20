#!/bin/bash
# first GPU round trip: diagnostics + SIMT-path parity + tensor-core GEMM parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
timeout 600 python tools/diag_gpu.py > gpurun_out/diag_main.log 2>&1; echo "diag_main exit $?"
timeout 300 python tools/diag_gpu.py --sweep-only > gpurun_out/diag_sweep.log 2>&1; echo "diag_sweep exit $?"
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "not tc_bf16" > gpurun_out/pytest_ops_simt.log 2>&1; echo "ops_simt exit $?"
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "tc_bf16 and (gemm or replay)" > gpurun_out/pytest_ops_tc.log 2>&1; echo "ops_tc exit $?"
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -k "fp32 or bf16_simt or bf16_tcgemm or attention_maps" > gpurun_out/pytest_model.log 2>&1; echo "model exit $?"
tail -5 gpurun_out/diag_main.log gpurun_out/pytest_ops_simt.log gpurun_out/pytest_ops_tc.log gpurun_out/pytest_model.log

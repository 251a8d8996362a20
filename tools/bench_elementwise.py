"""Achieved HBM GB/s of the bandwidth-bound kernels at the BASELINE configs[1] shapes (B200)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_vit_b200 import _lib as L, ops
DEV = "cuda"
B, N, D, HID = 256, 2049, 384, 1536
M = B * N
T = L.BF16
bf = torch.bfloat16

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

x = torch.randn(M, D, device=DEV); g = torch.randn(M, D, device=DEV)
y = torch.empty(M, D, dtype=bf, device=DEV); dy = torch.randn(M, D, device=DEV).to(bf)
w = torch.randn(D, device=DEV); b = torch.randn(D, device=DEV)
mean = torch.empty(M, device=DEV); rstd = torch.empty(M, device=DEV)
dx = torch.empty(M, D, device=DEV); gp = torch.empty(M, D, dtype=bf, device=DEV)
dw = torch.zeros(D, device=DEV); db = torch.zeros(D, device=DEV); cs = torch.zeros(D, device=DEV)
dh = torch.randn(M, HID, device=DEV).to(bf); cb = torch.zeros(HID, device=DEV)
xin = torch.randn(B, 8, 128, 256, device=DEV); cols = torch.empty(B * 2048, 128, dtype=bf, device=DEV)
rows = [
    ("ln_fwd", lambda: ops.ln_fwd(x, D, w, b, y, T, mean, rstd, M, D), M * D * 6),
    ("ln_bwd (+gp)", lambda: ops.ln_bwd(dy, T, x, D, mean, rstd, w, g, dx, D, dw, db, M, D, gp=gp, row_scale=None, rows_per_group=N, gp_colsum=cs), M * D * 16),
    ("ln_bwd", lambda: ops.ln_bwd(dy, T, x, D, mean, rstd, w, g, dx, D, dw, db, M, D), M * D * 14),
    ("ln_bwd (+gp, dropout)", lambda: ops.ln_bwd(dy, T, x, D, mean, rstd, w, g, dx, D, dw, db, M, D, gp=gp, row_scale=None, rows_per_group=N, drop=(1, 2, 0.1), gp_colsum=cs), M * D * 16),
    ("branch_grad_prep", lambda: ops.branch_grad_prep(g, M, D, None, N, None, gp, T, cs), M * D * 6),
    ("branch_grad_prep dropout", lambda: ops.branch_grad_prep(g, M, D, None, N, (1, 2, 0.1), gp, T, cs), M * D * 6),
    ("colsum [M,4D]", lambda: ops.colsum(dh, T, M, HID, HID, cb), M * HID * 2),
    ("im2col", lambda: ops.im2col(xin, cols, T, B, 8, 128, 256, 2, 8, 8), xin.numel() * 6),
]
peak = 6544.7
for name, fn, nbytes in rows:
    ms = timeit(fn)
    print(f"{name:28s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s  ({nbytes / ms / 1e6 / peak * 100:4.1f}% of measured HBM peak)", flush=True)

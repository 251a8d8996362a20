"""Launch one instance of every hot kernel at the BASELINE configs[1] shapes (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_vit_b200 import _lib as L  # noqa: E402
from neural_vit_b200 import ops  # noqa: E402

DEV = "cuda"
B, N, H, hd = (int(os.environ.get("PROF_B", "256")), 2049, 6, 64)
D, HID = H * hd, 4 * H * hd
M = B * N
E, T = L.ENGINE_TCGEN05, L.BF16
bf = torch.bfloat16
drop = (1234, 17, 0.1) if os.environ.get("PROF_DROPOUT", "0") == "1" else None


def r(*shape, dtype=bf, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(dtype)


x = r(M, D, dtype=torch.float32)
y = r(M, D)
w_qkv, w_proj, w_fc1, w_fc2 = r(3 * D, D, scale=0.05), r(D, D, scale=0.05), r(HID, D, scale=0.05), r(D, HID, scale=0.05)
bias_d, bias_3d, bias_h = r(D, dtype=torch.float32), r(3 * D, dtype=torch.float32), r(HID, dtype=torch.float32)
gamma = r(D, dtype=torch.float32)
qkv = torch.empty(M, 3 * D, dtype=bf, device=DEV)
ao = torch.empty(M, D, dtype=bf, device=DEV)
lse = torch.empty(B, H, N, device=DEV)
hpre, act = torch.empty(M, HID, dtype=bf, device=DEV), torch.empty(M, HID, dtype=bf, device=DEV)
hout = torch.empty(M, D, device=DEV)
mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
lnw, lnb = r(D, dtype=torch.float32), r(D, dtype=torch.float32)

ops.ln_fwd(x, D, lnw, lnb, y, T, mean, rstd, M, D)
ops.gemm(E, T, y, w_qkv, M, 3 * D, D, epilogue=L.EPI_STORE, out=qkv, bias=bias_3d)
ops.attn_fwd(E, T, qkv, ao, lse, B, N, H, hd, drop)
ops.gemm(E, T, ao, w_proj, M, D, D, epilogue=L.EPI_RESIDUAL, out=hout, bias=bias_d, resid=x, gamma=gamma, drop=drop)
ops.gemm(E, T, y, w_fc1, M, HID, D, epilogue=L.EPI_BIAS_GELU, out=act, aux=hpre, bias=bias_h, drop=drop)
ops.gemm(E, T, act, w_fc2, M, D, HID, epilogue=L.EPI_RESIDUAL, out=hout, bias=bias_d, resid=x, gamma=gamma, drop=drop)
# backward pieces
gp = torch.empty(M, D, dtype=bf, device=DEV)
cs = torch.zeros(D, device=DEV)
ops.branch_grad_prep(hout, M, D, None, N, drop, gp, T, cs)
dh = torch.empty(M, HID, dtype=bf, device=DEV)
ops.gemm(E, T, gp, w_fc2.T.contiguous(), M, HID, D, epilogue=L.EPI_GELU_BWD, out=dh, aux=hpre, drop=drop)
G = torch.zeros(D, HID, device=DEV)
ops.gemm(E, T, gp, act, D, HID, M, epilogue=L.EPI_ACCUM_F32, out=G, trans_a=True, trans_b=True)
cb = torch.zeros(HID, device=DEV)
ops.colsum(dh, T, M, HID, HID, cb)
dy = torch.empty(M, D, dtype=bf, device=DEV)
ops.gemm(E, T, dh, w_fc1.T.contiguous(), M, D, HID, epilogue=L.EPI_STORE, out=dy)
dx = torch.empty(M, D, device=DEV)
dw, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
ops.ln_bwd(dy, T, x, D, mean, rstd, lnw, hout, dx, D, dw, db, M, D, gp=gp, row_scale=None, rows_per_group=N, drop=drop,
           gp_colsum=cs)
dqkv = torch.empty_like(qkv)
ops.attn_bwd(E, T, qkv, ao, gp, lse, dqkv, B, N, H, hd, drop)
torch.cuda.synchronize()
print("prof_kernels done")

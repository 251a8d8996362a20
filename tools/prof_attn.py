import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_vit_b200 import _lib as L, ops
B, N, H, hd = int(os.environ.get("PROF_B", "16")), 2049, 6, 64
D = H * hd
mode = os.environ.get("PROF_DROPOUT", "0")   # "0", "1" or "both"
drops = {"0": [None], "1": [(1, 2, 0.1)], "both": [None, (1, 2, 0.1)]}[mode]
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
out = torch.empty(B * N, D, dtype=torch.bfloat16, device="cuda")
lse = torch.empty(B, H, N, device="cuda")
dout = torch.randn(B * N, D, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
use_kb = os.environ.get("PROF_KEEPBITS", "1") != "0"   # dropout keep-flag cache (the model's default path)
for _ in range(2):
    for drop in drops:
        kb = ops.attn_keepbits(L.ENGINE_TCGEN05, B, N, H, drop, qkv.device) if use_kb else None
        ops.attn_fwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, lse, B, N, H, hd, drop, keepbits=kb)
        ops.attn_bwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, dout, lse, dqkv, B, N, H, hd, drop, keepbits=kb)
torch.cuda.synchronize()
print("done")

"""Pipeline event trace of tc_attn_bwd (development aid): needs a library built with -DTVIT_ATTN_TRACE
(TVIT_LIB_PATH=ab/libtvit_trace.so python tools/attn_trace.py [--drop]).  Prints, for CTA (0,0,0), clock64 deltas of the
MMA warp / softmax warps 0 and 15 / drain warp 16 per query tile."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_vit_b200 import _lib as L  # noqa: E402
from neural_vit_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--drop", action="store_true")
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    B, N, D, H = a.batch, 2049, 384, 6
    M = B * N
    E, T = L.ENGINE_TCGEN05, L.BF16
    dev = "cuda"
    qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
    out = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, N, device=dev)
    dout = torch.randn(M, D, device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    d = (1234, 18, 0.1) if a.drop else None
    ops.attn_fwd(E, T, qkv, out, lse, B, N, H, 64, d)
    for _ in range(2):
        ops.attn_bwd(E, T, qkv, out, dout, lse, dqkv, B, N, H, 64, d)
    torch.cuda.synchronize()
    lib = L.load()
    n = 8 * 32 * 8
    buf = (ctypes.c_longlong * n)()
    lib.tvit_attn_bwd_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    rc = lib.tvit_attn_bwd_trace(buf, n)
    assert rc == 0, rc

    def ev(slot, i, e):
        return buf[(slot * 32 + i) * 8 + e]

    t0 = ev(1, 0, 0)
    print("fine (warp 0): half a: t(math done) | +tds_free +p_free +ds_free +STTM +STS || half b: t | +STS +wait::st +fence +arrive")
    for i in range(16):
        a = [ev(5, i, e) for e in range(6)]
        bq = [ev(6, i, e) for e in range(5)]
        print(f"{i:3d} | {a[0] - t0:7d} | " + " ".join(f"{a[k] - a[k - 1]:6d}" for k in range(1, 6)) +
              f" || {bq[0] - t0:7d} | " + " ".join(f"{bq[k] - bq[k - 1]:6d}" for k in range(1, 5)))
    print("tile |  MMA: SDnext pfull dqfree | smx0: a.sfull a.ld a.math a.waits b.sfull b.ld b.math b.waits | smx15 a.sfull "
          "b.waits | drain: dqfull dqfree done | MMA-B: pfull dVissued dKissued")
    for i in range(16):
        row = [ev(0, i, e) - t0 for e in range(3)]
        row += [ev(1, i, e) - t0 for e in range(8)]
        row += [ev(2, i, 0) - t0, ev(2, i, 7) - t0]
        row += [ev(3, i, e) - t0 for e in range(3)]
        row += [ev(4, i, e) - t0 for e in range(3)]
        print(f"{i:3d} | " + " ".join(f"{x:7d}" for x in row))


if __name__ == "__main__":
    main()

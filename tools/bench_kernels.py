"""Isolated timings of the hot-path kernels at the bench shapes (development aid; run under gpurun).

    python tools/bench_kernels.py [--batch 256] [--only gemm,attn,elem] [--reps 10]

Every op is timed with CUDA events on the launching stream over `reps` launches after 2 warm-up launches; operands
are far larger than the 126 MB L2 (M = batch * 2049 rows), so consecutive launches do not find their inputs cached.
Prints ms per launch and the TFLOP/s (GEMM / attention) or GB/s (elementwise) that corresponds to.
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_vit_b200 import _lib as L  # noqa: E402
from neural_vit_b200 import ops  # noqa: E402

DEV = "cuda"
E, T = L.ENGINE_TCGEN05, L.BF16


def timeit(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def bf(*shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).bfloat16()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--tokens", type=int, default=2049)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--only", default="gemm,attn,elem")
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    B, N, D = a.batch, a.tokens, a.dim
    H, hd, hid = D // 64, 64, 4 * D
    M = B * N
    only = set(a.only.split(","))
    drop = (1234, 18, 0.1)
    print(torch.cuda.get_device_name(0), f"B={B} N={N} D={D} M={M}", flush=True)

    def report(name, ms, flops=None, nbytes=None):
        extra = ""
        if flops:
            extra += f"  {flops / ms / 1e9:8.1f} TFLOP/s"
        if nbytes:
            extra += f"  {nbytes / ms / 1e6:8.1f} GB/s"
        print(f"{name:44s} {ms:8.3f} ms{extra}", flush=True)

    if "gemm" in only:
        y = bf(M, D)
        w1 = bf(hid, D, scale=1 / math.sqrt(D))
        b1 = torch.randn(hid, device=DEV)
        act, aux = torch.empty(M, hid, dtype=torch.bfloat16, device=DEV), torch.empty(M, hid, dtype=torch.bfloat16, device=DEV)
        fl = 2.0 * M * D * hid
        for d in (None, drop):
            ms = timeit(lambda: ops.gemm(E, T, y, w1, M, hid, D, epilogue=L.EPI_BIAS_GELU, out=act, aux=aux, bias=b1,
                                         drop=d), a.reps)
            report(f"fc1 BIAS_GELU {'drop' if d else 'nodrop'} [M,{hid}]x{D}", ms, fl, M * (D * 2 + hid * 4))
        ms = timeit(lambda: ops.gemm(E, T, y, w1, M, hid, D, epilogue=L.EPI_BIAS_GELU, out=act, bias=b1), a.reps)
        report(f"fc1 BIAS_GELU nodrop, no aux (inference forward)", ms, fl, M * (D * 2 + hid * 2))
        gp = bf(M, D)
        w2t = bf(hid, D, scale=1 / math.sqrt(D))
        dh = torch.empty(M, hid, dtype=torch.bfloat16, device=DEV)
        ms = timeit(lambda: ops.gemm(E, T, gp, w2t, M, hid, D, epilogue=L.EPI_GELU_BWD, out=dh, aux=aux), a.reps)
        report(f"GELU_BWD dgrad [M,{hid}]x{D}", ms, fl, M * (D * 2 + hid * 4))
        csf = torch.zeros(hid, device=DEV)
        ms = timeit(lambda: ops.gemm(E, T, gp, w2t, M, hid, D, epilogue=L.EPI_GELU_BWD, out=dh, aux=aux, colsum=csf), a.reps)
        report(f"GELU_BWD dgrad + fused colsum", ms, fl, M * (D * 2 + hid * 4))
        ms = timeit(lambda: ops.gemm(E, T, gp, w2t, M, hid, D, epilogue=L.EPI_STORE, out=dh), a.reps)
        report(f"STORE [M,{hid}]x{D} (same shape, plain)", ms, fl, M * (D * 2 + hid * 2))
        w2 = bf(D, hid, scale=1 / math.sqrt(hid))
        resid, hout = torch.randn(M, D, device=DEV), torch.empty(M, D, device=DEV)
        gamma, b2, rs = torch.randn(D, device=DEV), torch.randn(D, device=DEV), torch.ones(B, device=DEV)
        for d in (None, drop):
            ms = timeit(lambda: ops.gemm(E, T, act, w2, M, D, hid, epilogue=L.EPI_RESIDUAL, out=hout, bias=b2, resid=resid,
                                         gamma=gamma, row_scale=rs, rows_per_group=N, drop=d), a.reps)
            report(f"fc2 RESIDUAL {'drop' if d else 'nodrop'} [M,{D}]x{hid}", ms, fl, M * (hid * 2 + D * 8))
        wp = bf(D, D, scale=1 / math.sqrt(D))
        for d in (None, drop):
            ms = timeit(lambda: ops.gemm(E, T, y, wp, M, D, D, epilogue=L.EPI_RESIDUAL, out=hout, bias=b2, resid=resid,
                                         gamma=gamma, row_scale=rs, rows_per_group=N, drop=d), a.reps)
            report(f"proj RESIDUAL {'drop' if d else 'nodrop'} [M,{D}]x{D}", ms, 2.0 * M * D * D, M * (D * 2 + D * 8))
        wq = bf(3 * D, D, scale=1 / math.sqrt(D))
        qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=DEV)
        bq = torch.randn(3 * D, device=DEV)
        ms = timeit(lambda: ops.gemm(E, T, y, wq, M, 3 * D, D, epilogue=L.EPI_STORE, out=qkv, bias=bq), a.reps)
        report(f"qkv STORE [M,{3 * D}]x{D}", ms, 2.0 * M * D * 3 * D, M * (D * 2 + 3 * D * 2))
        dy = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
        ms = timeit(lambda: ops.gemm(E, T, dh, w2, M, D, hid, epilogue=L.EPI_STORE, out=dy), a.reps)
        report(f"dgrad STORE [M,{D}]x{hid}", ms, fl, M * (hid * 2 + D * 2))
        G = torch.zeros(hid, D, device=DEV)
        ms = timeit(lambda: ops.gemm(E, T, dh, y, hid, D, M, epilogue=L.EPI_ACCUM_F32, out=G, trans_a=True, trans_b=True),
                    a.reps)
        report(f"wgrad ACCUM [{hid},{D}]xM", ms, fl, M * (hid * 2 + D * 2))
        del act, aux, dh, resid, hout, qkv, dy

    if "attn" in only:
        qkv = bf(M, 3 * D)
        out = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
        lse = torch.empty(B, H, N, device=DEV)
        dout = bf(M, D)
        dqkv = torch.empty_like(qkv)
        fl = 4.0 * N * N * D * B
        for d in (None, drop):
            tag = "drop" if d else "nodrop"
            ms = timeit(lambda: ops.attn_fwd(E, T, qkv, out, lse, B, N, H, hd, d), a.reps)
            report(f"attn fwd {tag}", ms, fl)
            ms = timeit(lambda: ops.attn_bwd(E, T, qkv, out, dout, lse, dqkv, B, N, H, hd, d), a.reps)
            report(f"attn bwd {tag} (algorithmic 8 N^2 D)", ms, 2 * fl)
            csq = torch.zeros(3 * D, device=DEV)
            ms = timeit(lambda: ops.attn_bwd(E, T, qkv, out, dout, lse, dqkv, B, N, H, hd, d, colsum=csq), a.reps)
            report(f"attn bwd {tag} + fused qkv-bias colsum", ms, 2 * fl)
            if d:
                kb = ops.attn_keepbits(E, B, N, H, d, qkv.device)
                ms = timeit(lambda: ops.attn_fwd(E, T, qkv, out, lse, B, N, H, hd, d, keepbits=kb), a.reps)
                report(f"attn fwd {tag} + keep-bit cache written", ms, fl)
                ms = timeit(lambda: ops.attn_bwd(E, T, qkv, out, dout, lse, dqkv, B, N, H, hd, d, keepbits=kb), a.reps)
                report(f"attn bwd {tag}, masks from the keep-bit cache", ms, 2 * fl)
                ms = timeit(lambda: ops.attn_bwd(E, T, qkv, out, dout, lse, dqkv, B, N, H, hd, d, colsum=csq, keepbits=kb),
                            a.reps)
                report(f"attn bwd {tag}, keep-bit cache + fused colsum", ms, 2 * fl)
                del kb
            # again in the opposite order: under the power cap the clocks sag over a long run of launches, which would
            # otherwise always be charged to whichever variant is timed last
            ms = timeit(lambda: ops.attn_bwd(E, T, qkv, out, dout, lse, dqkv, B, N, H, hd, d), a.reps)
            report(f"attn bwd {tag} (again, after the colsum variant)", ms, 2 * fl)
        del qkv, out, dout, dqkv

    if "elem" in only:
        h = torch.randn(M, D, device=DEV)
        w, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
        yb = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
        mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
        ms = timeit(lambda: ops.ln_fwd(h, D, w, b, yb, T, mean, rstd, M, D), a.reps)
        report("ln_fwd", ms, None, M * D * 6)
        dyb = bf(M, D)
        gres, dx = torch.randn(M, D, device=DEV), torch.empty(M, D, device=DEV)
        dw, db, cs = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
        gpb = torch.empty(M, D, dtype=torch.bfloat16, device=DEV)
        rs = torch.ones(B, device=DEV)
        ms = timeit(lambda: ops.ln_bwd(dyb, T, h, D, mean, rstd, w, gres, dx, D, dw, db, M, D), a.reps)
        report("ln_bwd (plain)", ms, None, M * D * 14)
        for d in (None, drop):
            ms = timeit(lambda: ops.ln_bwd(dyb, T, h, D, mean, rstd, w, gres, dx, D, dw, db, M, D, gp=gpb, row_scale=rs,
                                           rows_per_group=N, drop=d, gp_colsum=cs), a.reps)
            report(f"ln_bwd + branch grad {'drop' if d else 'nodrop'}", ms, None, M * D * 16)
            ms = timeit(lambda: ops.branch_grad_prep(gres, M, D, rs, N, d, gpb, T, cs), a.reps)
            report(f"branch_grad_prep {'drop' if d else 'nodrop'}", ms, None, M * D * 6)
        big = bf(M, hid)
        csb = torch.zeros(hid, device=DEV)
        ms = timeit(lambda: ops.colsum(big, T, M, hid, hid, csb), a.reps)
        report(f"colsum [M,{hid}]", ms, None, M * hid * 2)
        big3 = bf(M, 3 * D)
        cs3 = torch.zeros(3 * D, device=DEV)
        ms = timeit(lambda: ops.colsum(big3, T, M, 3 * D, 3 * D, cs3), a.reps)
        report(f"colsum [M,{3 * D}]", ms, None, M * 3 * D * 2)


if __name__ == "__main__":
    main()

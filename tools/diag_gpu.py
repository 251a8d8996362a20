"""GPU diagnostics for the tensor-core kernels (development aid; run under gpurun).

Prints, for each GEMM layout, the relative error against an fp64 torch matmul and -- when it is
wrong -- a coarse error map, so that descriptor/swizzle mistakes can be told apart from pipeline
mistakes in ONE round trip.  For the MN-major (weight-gradient) layout it can sweep the UMMA
descriptor geometry through the TVIT_MN_DESC debug override.
"""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_vit_b200 import _lib as L  # noqa: E402
from neural_vit_b200 import ops  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def errmap(out, ref, rb=16, cb=32):
    d = (out.double() - ref.double()).abs()
    M, N = d.shape
    rows = []
    for r0 in range(0, min(M, 128), rb):
        rows.append(" ".join(f"{d[r0:r0 + rb, c0:c0 + cb].max().item():8.2e}" for c0 in range(0, min(N, 256), cb)))
    return "\n".join(rows)


def run(fn, what):
    try:
        r = fn()
        torch.cuda.synchronize()
        return r
    except Exception as e:  # noqa: BLE001
        print(f"!! {what}: {type(e).__name__}: {e}")
        return None


def gemm_nt(M, N, K, engine=L.ENGINE_TCGEN05):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(DEV).bfloat16()
    b = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(engine, L.BF16, a, b, M, N, K, epilogue=L.EPI_STORE, out=out)
    torch.cuda.synchronize()
    ref = a.double() @ b.double().T
    return out, ref


def gemm_tn(M, N, K, engine=L.ENGINE_TCGEN05):
    g = torch.Generator().manual_seed(M * 5 + N * 11 + K)
    a = torch.randn(K, M, generator=g).to(DEV).bfloat16()
    b = (torch.randn(K, N, generator=g) / math.sqrt(K)).to(DEV).bfloat16()
    out = torch.zeros(M, N, dtype=torch.float32, device=DEV)
    ops.gemm(engine, L.BF16, a, b, M, N, K, epilogue=L.EPI_ACCUM_F32, out=out, trans_a=True, trans_b=True)
    torch.cuda.synchronize()
    return out, a.double().T @ b.double()


def sweep():
    print("== MN-major descriptor sweep (lbo,sbo,kstep)")
    for lbo in (8192, 1024, 128, 0):
        for sbo in (1024, 8192, 128):
            for ks in (2048, 32, 256):
                os.environ["TVIT_MN_DESC"] = f"{lbo},{sbo},{ks}"
                r = run(lambda: gemm_tn(128, 128, 128), f"sweep {lbo},{sbo},{ks}")
                if r is not None:
                    print(f"  lbo={lbo:5d} sbo={sbo:5d} kstep={ks:5d} rel_err={rel(*r):.3e}", flush=True)
    os.environ.pop("TVIT_MN_DESC", None)


def main():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    L.require_device(0)
    if "--sweep-only" in sys.argv:
        sweep()
        return
    print("== K-major (NT) tcgen05 GEMM")
    for (M, N, K) in [(128, 128, 64), (128, 128, 128), (128, 256, 64), (256, 192, 256), (300, 384, 384),
                      (4098, 1152, 384), (1000, 1536, 384), (777, 384, 1536)]:
        r = run(lambda: gemm_nt(M, N, K), f"NT {M}x{N}x{K}")
        if r is None:
            continue
        e = rel(*r)
        print(f"NT {M:5d}x{N:5d}x{K:5d} rel_err={e:.3e} {'OK' if e < 1e-2 else 'BAD'}", flush=True)
        if e >= 1e-2:
            print(errmap(*r))
    print("== MN-major (TN, weight-gradient) tcgen05 GEMM")
    for (M, N, K) in [(128, 128, 64), (128, 128, 128), (128, 192, 64), (256, 256, 320), (384, 1536, 4098),
                      (1152, 384, 5000)]:
        r = run(lambda: gemm_tn(M, N, K), f"TN {M}x{N}x{K}")
        if r is None:
            continue
        e = rel(*r)
        print(f"TN {M:5d}x{N:5d}x{K:5d} rel_err={e:.3e} {'OK' if e < 1e-3 else 'BAD'}", flush=True)
        if e >= 1e-3 and M <= 256:
            print(errmap(*r))

    print("== timing (CUDA events, 20 iters after 5 warm-up)")
    def timeit(fn, iters=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters
    for (M, N, K) in [(524544, 1152, 384), (524544, 384, 384), (524544, 1536, 384), (524544, 384, 1536)]:
        def mk():
            a = torch.randn(M, K, device=DEV).bfloat16()
            b = torch.randn(N, K, device=DEV).bfloat16()
            out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
            return a, b, out
        r = run(mk, "alloc")
        if r is None:
            continue
        a, b, out = r
        t = run(lambda: timeit(lambda: ops.gemm(L.ENGINE_TCGEN05, L.BF16, a, b, M, N, K, epilogue=L.EPI_STORE, out=out)), "time NT")
        tt = run(lambda: timeit(lambda: torch.matmul(a, b.T, out=out)), "time torch")
        if t and tt:
            fl = 2.0 * M * N * K
            print(f"NT {M}x{N}x{K}: ours {t:.3f} ms ({fl / t / 1e9:.0f} TFLOP/s)  cuBLAS {tt:.3f} ms ({fl / tt / 1e9:.0f} TFLOP/s)", flush=True)
        del a, b, out
    for (M, N, K) in [(1536, 384, 524544), (384, 384, 524544)]:
        a = torch.randn(K, M, device=DEV).bfloat16()
        b = torch.randn(K, N, device=DEV).bfloat16()
        out = torch.zeros(M, N, device=DEV)
        t = run(lambda: timeit(lambda: ops.gemm(L.ENGINE_TCGEN05, L.BF16, a, b, M, N, K, epilogue=L.EPI_ACCUM_F32, out=out, trans_a=True, trans_b=True)), "time TN")
        tt = run(lambda: timeit(lambda: torch.matmul(a.T, b)), "time torch TN")
        if t and tt:
            fl = 2.0 * M * N * K
            print(f"TN {M}x{N}x{K}: ours {t:.3f} ms ({fl / t / 1e9:.0f} TFLOP/s)  cuBLAS {tt:.3f} ms ({fl / tt / 1e9:.0f} TFLOP/s)", flush=True)
        del a, b, out


if __name__ == "__main__":
    main()

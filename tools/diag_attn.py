"""GPU diagnostics for the tcgen05 attention kernels (development aid; run under gpurun)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_vit_b200 import _lib as L  # noqa: E402
from neural_vit_b200 import ops  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def ref_attn(qkv, B, N, H, hd):
    D = H * hd
    t = qkv.double().reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (t[0] @ t[1].transpose(-2, -1)) * hd ** -0.5
    return (torch.softmax(s, -1) @ t[2]).transpose(1, 2).reshape(B * N, D), torch.logsumexp(s, -1)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    do_bwd = "--bwd" in sys.argv
    hd = 64
    for (B, N, H) in [(1, 128, 1), (1, 256, 1), (2, 17, 1), (2, 257, 3), (1, 2049, 2), (3, 130, 6)]:
        D = H * hd
        g = torch.Generator().manual_seed(B * 1000 + N)
        qkv = torch.randn(B * N, 3 * D, generator=g).to(DEV).bfloat16()
        out = torch.empty(B * N, D, dtype=torch.bfloat16, device=DEV)
        lse = torch.empty(B, H, N, device=DEV)
        try:
            ops.attn_fwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, lse, B, N, H, hd)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"!! fwd B={B} N={N} H={H}: {e}")
            return
        qd = qkv.double().requires_grad_(True)
        ro, rl = ref_attn(qd, B, N, H, hd)
        print(f"fwd B={B} N={N:5d} H={H}: out rel_err={rel(out, ro):.3e} lse rel_err={rel(lse, rl):.3e}", flush=True)
        if rel(out, ro) > 2e-2:
            d = (out.double() - ro).abs().reshape(B, N, H, hd)
            print("  err by row block (b=0,h=0):", [f"{d[0, i:i + 32, 0].max().item():.2e}" for i in range(0, min(N, 256), 32)])
            print("  err by col block (b=0,h=0):", [f"{d[0, :, 0, i:i + 8].max().item():.2e}" for i in range(0, 64, 8)])
        if do_bwd:
            dout = torch.randn(B * N, D, generator=g).to(DEV).bfloat16()
            ro.backward(dout.double())
            dqkv = torch.empty_like(qkv)
            try:
                ops.attn_bwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, dout, lse, dqkv, B, N, H, hd)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f"!! bwd B={B} N={N} H={H}: {e}")
                return
            gq = qd.grad.reshape(B * N, 3, D)
            got = dqkv.reshape(B * N, 3, D)
            errs = [rel(got[:, i], gq[:, i]) for i in range(3)]
            print(f"bwd B={B} N={N:5d} H={H}: dq {errs[0]:.3e} dk {errs[1]:.3e} dv {errs[2]:.3e}", flush=True)
            for i, nm in enumerate("qkv"):
                if errs[i] > 3e-2:
                    d = (got[:, i].double() - gq[:, i]).abs().reshape(B, N, H, hd)
                    print(f"  d{nm} err by row block:", [f"{d[0, r:r + 32, 0].max().item():.2e}" for r in range(0, min(N, 256), 32)])
    # box calibration: cuBLAS bf16 GEMM (boxes differ by a few % in sustained clocks)
    a = torch.randn(8192, 8192, device=DEV).bfloat16()
    bb = torch.randn(8192, 8192, device=DEV).bfloat16()
    t = timeit(lambda: torch.matmul(a, bb), iters=20)
    print(f"calibration cuBLAS 8192^3 bf16: {2 * 8192**3 / t / 1e9:.0f} TFLOP/s", flush=True)
    del a, bb
    # timing at the C2 shape
    B, N, H = 256, 2049, 6
    D = H * hd
    qkv = torch.randn(B * N, 3 * D, device=DEV).bfloat16()
    out = torch.empty(B * N, D, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B, H, N, device=DEV)
    t = timeit(lambda: ops.attn_fwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, lse, B, N, H, hd))
    fl = 4.0 * N * N * D * B
    print(f"fwd C2 shape: {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s", flush=True)
    t = timeit(lambda: ops.attn_fwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, lse, B, N, H, hd, (1, 2, 0.1)))
    print(f"fwd C2 shape dropout 0.1: {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s", flush=True)
    if do_bwd:
        dout = torch.randn(B * N, D, device=DEV).bfloat16()
        dqkv = torch.empty_like(qkv)
        t = timeit(lambda: ops.attn_bwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, dout, lse, dqkv, B, N, H, hd))
        print(f"bwd C2 shape: {t:.3f} ms  {2.5 * fl / t / 1e9:.0f} TFLOP/s", flush=True)
        t = timeit(lambda: ops.attn_bwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, dout, lse, dqkv, B, N, H, hd, (1, 2, 0.1)))
        print(f"bwd C2 shape dropout 0.1: {t:.3f} ms  {2.5 * fl / t / 1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()

"""Inference (evaluate()-style) forward throughput of the B200 model on one GPU: eval mode, no grad, bf16.
    python tools/bench_eval.py [--batch 256] [--steps 10]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_vit_b200 as nv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=8)
    ap.add_argument("--time", type=int, default=256)
    ap.add_argument("--embed-dim", type=int, default=384)
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    # BASELINE configs[1] geometry by default (bench.py's c2): 8 x 128 x 256 volume, D384 / H6 / L8, N = 2049 tokens
    cfg = nv.Temporal3DViTConfig(n_trials=a.trials, freq_size=128, time_size=a.time, embed_dim=a.embed_dim,
                                 n_heads=a.embed_dim // 64, n_layers=a.layers)
    m = nv.Temporal3DViT(cfg, precision="bf16").cuda().eval()
    x = torch.randn(a.batch, cfg.n_trials, cfg.freq_size, cfg.time_size, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"eval forward D{a.embed_dim}/L{a.layers} {a.trials}x128x{a.time} batch {a.batch}: {ms:.2f} ms/step, "
          f"{a.batch / ms * 1e3:.1f} samples/s")


if __name__ == "__main__":
    main()

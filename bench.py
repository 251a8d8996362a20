#!/usr/bin/env python
"""Benchmark of the Temporal 3D ViT training hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--dropout p]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one synthetic batch exactly as the reference loop drives it
(train.py:223-227): zero_grad, forward, class-weighted CrossEntropy, backward (+ bucketed gradient
all-reduce at N > 1) and the AdamW update.  Metric: train samples/sec (BASELINE.json), whole job.

Workload (N = 1): BASELINE.json configs[1] -- default "small" Temporal 3D ViT (D384/H6/L8), input
8 x 128 x 256 (N = 2049 tokens), batch 256 per GPU, bf16 tensor-core path, reference-default dropout
rates (0.1/0.1/0.1), synthetic WT/FMR1 labels.  N > 1 keeps the per-GPU batch (weak scaling).

One JSON line on stdout (rank 0).  `value` has inputs resident in HBM; `e2e` goes through the public
module call with pinned host inputs copied H2D and the loss read back D2H inside the timed region.
`--impl reference` times the oracle port of the reference (PyTorch fp32 on the host cores) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train samples/sec fwd+bwd"
UNIT = "samples/s"
WORKLOAD = "small D384/H6/L8, 8x128x256 (N=2049 tokens), batch 256 per GPU, bf16, dropout 0.1/0.1/0.1"


def flops_per_sample(cfg, with_bwd=True):
    """Algorithmic FLOPs: fwd = 2 n P D + L (24 N D^2 + 4 N^2 D) + 2 D^2 + 4 D;
    fwd+bwd = 3x except the patch embed (no input gradient) = 2x."""
    n, P, D, L = cfg.n_patches, cfg.patch_dim, cfg.embed_dim, cfg.n_layers
    N = n + 1
    hid_ratio = cfg.mlp_ratio
    embed = 2.0 * n * P * D
    gemm = L * (2.0 * N * D * 3 * D + 2.0 * N * D * D + 4.0 * N * D * D * hid_ratio)
    attn = L * 4.0 * N * N * D
    head = 2.0 * D * D + 2.0 * D * cfg.n_classes
    fwd = embed + gemm + attn + head
    if not with_bwd:
        return fwd
    return 2.0 * embed + 3.0 * (gemm + attn + head)


def attn_bwd_flops(cfg, batch):
    """Algorithmic FLOPs of one attention-backward launch: 2x the forward's 4 N^2 D per sample."""
    N, D = cfg.n_patches + 1, cfg.embed_dim
    return 2.0 * 4.0 * N * N * D * batch


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic_bytes(batch, args):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full`
    capture of the final build (profiles/r1_ncu_attn_final_B256_dropout.json, written by tools/ncu_summary.py); only
    valid for the shape it was captured on."""
    if (batch, args.layers, args.embed_dim, args.trials, args.time) != (256, 8, 384, 8, 256):
        return None
    path = os.path.join(ROOT, "profiles", "r1_ncu_attn_final_B256_dropout.json")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    try:
        with open(path) as fh:
            k = next(e for e in json.load(fh) if "tc_attn_bwd_kernel<1>" in e["kernel"])
        total = 0.0
        for m in ("dram_rd", "dram_wr"):
            v, u = k[m].split()
            total += float(v) * scale[u]
        return total
    except (OSError, KeyError, ValueError, StopIteration):
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return p.get("bf16_tflops_sustained", 1351.4), p.get("bf16_tflops", 1593.7), "measured"
    return 1400.0, 1590.0, "fallback"


# ---------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_steps(cfg_kwargs, sample_batch, steps, warmup, seed=0):
    """Time `steps` fwd+CE+bwd steps of the oracle (fp32, train mode with the reference-default dropout
    rates, masks drawn inside the timed region) on a batch of `sample_batch`.  Returns seconds per step."""
    from oracle import vit_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(**cfg_kwargs)
    params = O.random_params(cfg, seed=1234)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(sample_batch, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g)
    y = torch.randint(0, 2, (sample_batch,), generator=g)
    cw = torch.tensor([0.8, 1.3])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        masks = O.draw_masks(cfg, sample_batch, generator=g)
        O.loss_and_grads(x, y, params, cfg, class_weight=cw, label_smoothing=0.05, masks=masks)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return statistics.median(times), cores


def run_reference(args, cfg_kwargs):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = 1
    sec, cores = cpu_reference_steps(cfg_kwargs, sample, args.steps, args.warmup)
    value = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "host": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} sample(s) of the batch-256 step per timed step (same per-sample shapes, "
                                   "train mode, dropout masks drawn in the timed region)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--dropout", type=float, default=None, help="override all three dropout rates (default: 0.1)")
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--embed-dim", type=int, default=384)
    ap.add_argument("--heads", type=int, default=6)
    ap.add_argument("--trials", type=int, default=8)
    ap.add_argument("--freq", type=int, default=128)
    ap.add_argument("--time", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print a per-op time breakdown to stderr")
    args = ap.parse_args()

    cfg_kwargs = dict(n_trials=args.trials, freq_size=args.freq, time_size=args.time, embed_dim=args.embed_dim,
                      n_heads=args.heads, n_layers=args.layers)
    if args.dropout is not None:
        cfg_kwargs.update(dropout=args.dropout, attention_dropout=args.dropout, drop_path=args.dropout)
    if args.impl == "reference":
        return run_reference(args, cfg_kwargs)

    import torch.distributed as dist
    import neural_vit_b200 as nv
    from neural_vit_b200 import ops
    from neural_vit_b200.ddp import BucketedAllReduce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    cfg = nv.Temporal3DViTConfig(**cfg_kwargs)
    torch.manual_seed(1234)
    model = nv.Temporal3DViT(cfg, precision="bf16").to(dev)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=0.01)
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([0.8, 1.3], device=dev), label_smoothing=0.05)
    ddp = BucketedAllReduce(model) if world > 1 else None

    B = args.batch
    g = torch.Generator().manual_seed(rank)
    y_host = torch.randint(0, 2, (B,), generator=g)
    x_host = (torch.randn(B, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g)
              + 0.5 * y_host[:, None, None, None].float()).pin_memory()
    y_host = y_host.pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        logits = model(x)
        loss = crit(logits, y)
        loss.backward()
        if ddp is not None:
            ddp.finish()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(get_batch, read_loss):
        for _ in range(args.warmup):
            loss = step(*get_batch())
            if read_loss:
                loss.item()
        barrier()
        ops.LAUNCHES["count"] = 0
        ops.TIMED = {"attn_bwd": []}
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            loss = step(*get_batch())
            if read_loss:
                loss.item()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        timed = ops.TIMED
        ops.TIMED = None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, ops.LAUNCHES["count"], timed

    # ---- resident-input measurement (value) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches, timed = timed_loop(lambda: (x_dev, y_dev), read_loss=False)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    attn_ms = [s.elapsed_time(e) for s, e in timed["attn_bwd"]]
    attn_ms_avg = sum(attn_ms) / max(len(attn_ms), 1)

    # ---- end-to-end measurement: pinned host inputs, H2D inside the timed region, loss read back ----
    # The input pipeline is the usual double-buffered one: the copy of step k+1's batch from pinned host memory runs
    # on a side stream while step k computes.  All K copies and all K loss read-backs lie inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [(torch.empty_like(x_dev), torch.empty_like(y_dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(k):
        with torch.cuda.stream(copy_stream):
            bufs[k % 2][0].copy_(x_host, non_blocking=True)
            bufs[k % 2][1].copy_(y_host, non_blocking=True)
            ready[k % 2].record(copy_stream)

    def e2e_steps(n):
        issue_copy(0)
        for k in range(n):
            torch.cuda.current_stream().wait_event(ready[k % 2])
            if k + 1 < n:
                issue_copy(k + 1)  # buffer (k+1) % 2 was last read by step k-1, which loss.item() has synchronised
            step(*bufs[k % 2]).item()

    e2e_steps(args.warmup)
    barrier()
    s_e, e_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_e.record()
    e2e_steps(args.steps)
    e_e.record()
    barrier()
    ms_e2e = s_e.elapsed_time(e_e)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)

    breakdown = None
    if args.breakdown and rank == 0:
        names = ["attn_fwd", "attn_bwd"] + [f"gemm_e{i}" for i in range(6)] + ["gemm_e4_tn"]
        ops.TIMED = {n: [] for n in names}
        if os.environ.get("TVIT_BENCH_DETAIL"):
            ops.TIMED = {"detail": []}
        step(x_dev, y_dev)
        torch.cuda.synchronize()
        breakdown = {n: (round(sum(s.elapsed_time(e) for s, e in v), 3), len(v)) if "detail" in ops.TIMED
                     else round(sum(s.elapsed_time(e) for s, e in v), 3) for n, v in ops.TIMED.items() if v}
        ops.TIMED = None
        print("per-op ms in one step:", json.dumps(breakdown), file=sys.stderr)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    sustained, burst, which = load_peaks()
    fl_step = flops_per_sample(cfg) * B            # per GPU
    model_tflops = fl_step / (ms_step * 1e-3) / 1e12
    attn_tflops = attn_bwd_flops(cfg, B) / (attn_ms_avg * 1e-3) / 1e12 if attn_ms_avg > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD if (B, args.layers, args.embed_dim, args.dropout, args.trials, args.time) ==
                   (256, 8, 384, None, 8, 256)
                   else f"D{args.embed_dim}/H{args.heads}/L{args.layers}, {args.trials}x{args.freq}x{args.time}, "
                        f"batch {B} per GPU, bf16, dropout {cfg.dropout}",
                   "step": "zero_grad + forward + weighted CE + backward" + (" + bucketed NCCL all-reduce" if world > 1 else "")
                           + " + AdamW", "global_batch": world * B, "tokens_per_sample": cfg.n_patches + 1,
                   "parallelism": f"dp{world}", "l2": "per-step working set (>= 268 MB input, tens of GB of activations) exceeds the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "model_tflops_per_gpu": model_tflops,
        "model_frac_of_peak": model_tflops / sustained,
        "roofline": {"bound": "tensor", "kernel": "tc_attn_bwd_kernel (+prep/finish)", "achieved": attn_tflops,
                     "peak": sustained, "peak_kind": f"{which} sustained cuBLAS bf16", "unit": "TFLOP/s",
                     "frac": attn_tflops / sustained, "ms_per_launch": attn_ms_avg, "launches_timed": len(attn_ms),
                     "traffic": ncu_traffic_bytes(B, args)},
    }
    if breakdown:
        line["breakdown_ms"] = breakdown
    if world == 1 and not args.no_cpu_baseline:
        sample = 1
        sec, cores = cpu_reference_steps(cfg_kwargs, sample, steps=2, warmup=1)
        line["cpu_baseline"] = {"value": sample / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{sample} sample(s) per step of the same per-sample shapes, fp32 oracle, train "
                                          "mode with dropout masks drawn in the timed region; median of 2 steps"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

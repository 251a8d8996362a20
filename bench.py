#!/usr/bin/env python
"""Benchmark of the Temporal 3D ViT training hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--impl reference] ...
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one synthetic batch exactly as the reference loop drives it
(train.py:223-227): zero_grad, forward, class-weighted label-smoothed CrossEntropy, backward (+ bucketed gradient
all-reduce at N > 1) and the AdamW update.  Metric: train samples/sec (BASELINE.json), whole job.

Workloads (BASELINE.json configs; --config):
  c2 (default)  small D384/H6/L8, input 8 x 128 x 256 (N = 2049 tokens), batch 256 per GPU        <- the headline
  c3            depth-12 variant of c2, batch 256 per GPU
  c4            long sequence 32 x 128 x 512 (N = 16385 tokens), small model, batch 8 per GPU
  c5            large D768/H12/L24, batch 64 per GPU
  c1            c2's model at batch 8 (the reference's CPU-runnable case; used by the CPU legs)
All in bf16 on the tensor-core path with the reference-default dropout rates (0.1/0.1/0.1) and synthetic WT/FMR1
labels; N > 1 keeps the per-GPU batch (weak scaling).

One JSON line on stdout (rank 0).  `value` has inputs resident in HBM; `e2e` goes through the package's public API
(DevicePrefetcher -> Temporal3DViT -> CrossEntropyLoss -> FusedAdamW) with pinned host inputs copied H2D and the loss
read back D2H every step inside the timed region.  Extra objects: `roofline` (dominant kernel, CUDA-event timed
inside the step), `cpu_baseline` (the reference module on the host cores, bounded sample), `gpu_eager_baseline`
(the reference arithmetic in CUDA eager on the same GPU: the kernel-level comparator).
`--impl reference` times the reference's own CPU implementation (the unmodified reference module from baseline/_ref
when installed, else the oracle port) on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import importlib.util
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

CONFIGS = {
    "c1": dict(trials=8, freq=128, time=256, embed_dim=384, heads=6, layers=8, batch=8),
    "c2": dict(trials=8, freq=128, time=256, embed_dim=384, heads=6, layers=8, batch=256),
    "c3": dict(trials=8, freq=128, time=256, embed_dim=384, heads=6, layers=12, batch=256),
    "c4": dict(trials=32, freq=128, time=512, embed_dim=384, heads=6, layers=8, batch=8),
    "c5": dict(trials=8, freq=128, time=256, embed_dim=768, heads=12, layers=24, batch=64),
}
CONFIG_NAMES = {
    "c1": "BASELINE configs[0]: small, batch 8 (CPU case)", "c2": "BASELINE configs[1]: small, batch 256",
    "c3": "BASELINE configs[2]: depth 12", "c4": "BASELINE configs[3]: long sequence",
    "c5": "BASELINE configs[4]: large D768/H12/L24",
}


def workload_string(name, a, dropout):
    n_tok = (a.trials // 2) * (a.freq // 8) * (a.time // 8) + 1
    return (f"{name} ({CONFIG_NAMES.get(name, 'custom')}): D{a.embed_dim}/H{a.heads}/L{a.layers}, "
            f"{a.trials}x{a.freq}x{a.time} (N={n_tok} tokens), batch {a.batch} per GPU, bf16, "
            f"dropout {dropout}/{dropout}/{dropout}")


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


def ncu_traffic_bytes(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full`
    capture (written by tools/ncu_summary.py); only valid for the c2 shape it was captured on."""
    if name != "c2":
        return None
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for fn in ("r2b_ncu_attn_keepbits_B256_dropout.json", "r2_ncu_attn_B256_dropout.json",
               "r1_ncu_attn_final_B256_dropout.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as fh:
                k = next(e for e in json.load(fh) if "tc_attn_bwd_kernel" in e["kernel"])
            total = 0.0
            for m in ("dram_rd", "dram_wr"):
                v, u = k[m].split()
                total += float(v) * scale[u]
            return total
        except (OSError, KeyError, ValueError, StopIteration):
            continue
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return p.get("bf16_tflops_sustained", 1351.4), p.get("bf16_tflops", 1593.7), "measured"
    return 1400.0, 1590.0, "fallback"


# ---------------------------------------------------------------------------------------------------------
# the reference's own implementation (checker side): unmodified module from baseline/_ref, else the oracle port
# ---------------------------------------------------------------------------------------------------------
def load_reference_module():
    path = os.path.join(ROOT, "baseline", "_ref", "temporal_vit", "models", "model.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("tvit_reference_model", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["tvit_reference_model"] = mod     # dataclasses resolves cls.__module__ while decorating the config
    spec.loader.exec_module(mod)
    return mod


def model_kwargs(a, dropout):
    kw = dict(n_trials=a.trials, freq_size=a.freq, time_size=a.time, embed_dim=a.embed_dim, n_heads=a.heads,
              n_layers=a.layers)
    if dropout is not None:
        kw.update(dropout=dropout, attention_dropout=dropout, drop_path=dropout)
    return kw


def make_reference_step(cfg_kwargs, device, autocast_dtype=None, seed=1234):
    """Returns (step(x, y) -> loss, kind): zero_grad + fwd + weighted label-smoothed CE + bwd + AdamW of the reference
    in train mode with its own nn.Dropout / DropPath (model.py:57-71,102-117; train.py:223-227).  kind "reference" =
    the unmodified reference module; "port" (only when baseline/_ref is absent) = the oracle restatement, fwd + CE +
    bwd with masks drawn in the timed region."""
    ref = load_reference_module()
    cw = torch.tensor([0.8, 1.3], device=device)
    if ref is not None:
        torch.manual_seed(seed)
        model = ref.Temporal3DViT(ref.Temporal3DViTConfig(**cfg_kwargs)).to(device).train()
        crit = torch.nn.CrossEntropyLoss(weight=cw, label_smoothing=0.05)
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=0.01)      # train.py:154-156

        def step(x, y):                                                               # train.py:223-227
            opt.zero_grad()
            if autocast_dtype is not None:
                with torch.autocast(device_type=torch.device(device).type, dtype=autocast_dtype):
                    loss = crit(model(x).float(), y)
            else:
                loss = crit(model(x), y)
            loss.backward()
            opt.step()
            return loss
        return step, "reference"
    from oracle import vit_oracle as O
    cfg = O.OracleConfig(**cfg_kwargs)
    params = O.random_params(cfg, seed=seed, device=device)
    g = torch.Generator(device=device).manual_seed(seed)

    def step(x, y):
        masks = O.draw_masks(cfg, x.shape[0], generator=g, device=device)
        if autocast_dtype is not None:
            with torch.autocast(device_type=torch.device(device).type, dtype=autocast_dtype):
                return O.loss_and_grads(x, y, params, cfg, class_weight=cw, label_smoothing=0.05, masks=masks)[1]
        return O.loss_and_grads(x, y, params, cfg, class_weight=cw, label_smoothing=0.05, masks=masks)[1]
    return step, "port"


def cpu_reference_steps(cfg_kwargs, sample_batch, steps, warmup, seed=0):
    """Time `steps` fwd+CE+bwd steps of the reference on the host cores on a batch of `sample_batch`.
    Returns (seconds per step, cores, kind)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = make_reference_step(cfg_kwargs, "cpu")
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(sample_batch, cfg_kwargs["n_trials"], cfg_kwargs["freq_size"], cfg_kwargs["time_size"], generator=g)
    y = torch.randint(0, 2, (sample_batch,), generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(x, y)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return statistics.median(times), cores, kind


def gpu_eager_baseline(cfg_kwargs, dev, steps=3):
    """The reference arithmetic in CUDA eager on this GPU (SURVEY.md section 8d "kernel to beat"): fp32 as train.py runs
    it and bf16 autocast, at the largest batch that fits (the (B,H,N,N) score tensors of model.py:111-113 are saved for
    backward: ~2 GB per sample at the c2 shape)."""
    out = {}
    for tag, dt in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        for batch in (64, 32, 16, 8, 4, 2, 1):
            try:
                step, kind = make_reference_step(cfg_kwargs, dev, autocast_dtype=dt)
                x = torch.randn(batch, cfg_kwargs["n_trials"], cfg_kwargs["freq_size"], cfg_kwargs["time_size"],
                                device=dev)
                y = torch.randint(0, 2, (batch,), device=dev)
                step(x, y)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(steps):
                    step(x, y)
                e.record()
                torch.cuda.synchronize()
                ms = s.elapsed_time(e) / steps
                out[tag] = {"value": batch / (ms * 1e-3), "unit": UNIT, "batch": batch, "ms_per_step": ms,
                            "kind": kind}
                del step, x, y
                torch.cuda.empty_cache()
                break
            except torch.OutOfMemoryError:
                step = x = y = None
                torch.cuda.empty_cache()
                continue
    out["note"] = ("reference module (model.py) zero_grad + fwd + CE + bwd + AdamW in train mode, PyTorch eager on the "
                   "same GPU, largest power-of-two batch that fits")
    return out


def run_reference(args, cfg_kwargs, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = args.cpu_sample
    sec, cores, kind = cpu_reference_steps(cfg_kwargs, sample, args.steps, args.warmup)
    value = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "host": "cpu",
                   "step": "zero_grad + forward + weighted CE + backward + AdamW of the reference on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} sample(s) of the workload's per-sample shape per timed step (the full "
                                   f"per-GPU batch would take minutes per step on {cores} cores), fp32, train mode with "
                                   "the reference's own nn.Dropout / DropPath inside the timed region; median step"},
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--dropout", type=float, default=None, help="override all three dropout rates (default: 0.1)")
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--embed-dim", type=int, default=None)
    ap.add_argument("--heads", type=int, default=None)
    ap.add_argument("--trials", type=int, default=None)
    ap.add_argument("--freq", type=int, default=None)
    ap.add_argument("--time", type=int, default=None)
    ap.add_argument("--stock-loop", action="store_true",
                    help="drive the step with torch.optim.AdamW + torch CrossEntropyLoss (the reference's unchanged "
                         "loop) instead of FusedAdamW + the device loss")
    ap.add_argument("--cpu-sample", type=int, default=1, help="samples per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print a per-op time breakdown to stderr")
    args = ap.parse_args()
    custom = False
    for k, v in CONFIGS[args.config].items():
        if getattr(args, k) is None:
            setattr(args, k, v)
        else:
            custom = custom or getattr(args, k) != v
    name = args.config if not custom else f"{args.config}*"
    cfg_kwargs = model_kwargs(args, args.dropout)
    drop_str = 0.1 if args.dropout is None else args.dropout
    workload = workload_string(name, args, drop_str)
    if args.impl == "reference":
        return run_reference(args, cfg_kwargs, workload)

    import torch.distributed as dist
    import neural_vit_b200 as nv
    from neural_vit_b200 import ops

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
    cw = torch.tensor([0.8, 1.3], device=dev)
    ddp = nv.BucketedAllReduce(model, measure_exposed=True) if world > 1 else None
    if args.stock_loop:
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=0.01)
        crit = torch.nn.CrossEntropyLoss(weight=cw, label_smoothing=0.05)
    else:
        opt = nv.FusedAdamW(model.parameters(), lr=3e-4, weight_decay=0.01, model=model)
        crit = nv.CrossEntropyLoss(weight=cw, label_smoothing=0.05, metrics=nv.DeviceMetrics(dev))

    B = args.batch
    g = torch.Generator().manual_seed(rank)
    y_host = torch.randint(0, 2, (B,), generator=g)
    x_host = (torch.randn(B, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g)
              + 0.5 * y_host[:, None, None, None].float()).pin_memory()
    y_host = y_host.pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step(x, y):
        opt.zero_grad()
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

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # ---- resident-input measurement (value) ----
    for _ in range(args.warmup):
        step(x_dev, y_dev)
    barrier()
    if ddp is not None:
        ddp.exposed_events.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.LAUNCHES["count"] = 0
    ops.TIMED = {"attn_bwd": []}
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        step(x_dev, y_dev)
    e.record()
    barrier()
    ms_total = max_over_ranks(s.elapsed_time(e))
    launches, timed = ops.LAUNCHES["count"], ops.TIMED
    ops.TIMED = None
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    attn_ms = [a.elapsed_time(b) for a, b in timed["attn_bwd"]]
    attn_ms_avg = sum(attn_ms) / max(len(attn_ms), 1)
    exposed_ms = ddp.exposed_ms() if ddp is not None else 0.0
    if hasattr(crit, "metrics") and crit.metrics is not None:
        crit.metrics.reset()

    # ---- end-to-end measurement through the package's public API: the input feed (DevicePrefetcher: pinned host
    # batch -> async H2D on a side stream, double buffered) + model + loss + optimizer, loss read back every step ----
    def host_batches(n):
        for _ in range(n):
            yield x_host, y_host

    def e2e_steps(n):
        feed = nv.DevicePrefetcher(host_batches(n), dev)
        for xb, yb in feed:
            step(xb, yb).item()
        return feed.h2d_bytes

    e2e_steps(args.warmup)
    barrier()
    s_e, e_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_e.record()
    h2d = e2e_steps(args.steps)
    e_e.record()
    barrier()
    ms_e2e = max_over_ranks(s_e.elapsed_time(e_e))
    e2e_value = world * B / (ms_e2e / args.steps * 1e-3)

    breakdown = None
    if args.breakdown and rank == 0:
        names = ["attn_fwd", "attn_bwd"] + [f"gemm_e{i}" for i in range(7)] + ["gemm_e4_tn"]
        ops.TIMED = {n: [] for n in names}
        if os.environ.get("TVIT_BENCH_DETAIL"):
            ops.TIMED = {"detail": []}
        step(x_dev, y_dev)
        torch.cuda.synchronize()
        breakdown = {n: (round(sum(a.elapsed_time(b) for a, b in v), 3), len(v)) if "detail" in ops.TIMED
                     else round(sum(a.elapsed_time(b) for a, b in v), 3) for n, v in ops.TIMED.items() if v}
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
    step_desc = ("zero_grad + forward + weighted CE + backward" + (" + bucketed NCCL all-reduce (AVG, grads written "
                 "straight into the buckets)" if world > 1 else "") + " + AdamW; "
                 + ("torch.optim.AdamW + torch CE (stock loop)" if args.stock_loop
                    else "FusedAdamW + device CE/metrics (this package)"))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload, "step": step_desc, "global_batch": world * B,
                   "tokens_per_sample": cfg.n_patches + 1, "parallelism": f"dp{world}",
                   "l2": "per-step working set (>= 268 MB input, tens of GB of activations) exceeds the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "model_tflops_per_gpu": model_tflops,
        "model_frac_of_peak": model_tflops / sustained,
        "roofline": {"bound": "tensor", "kernel": "tc_attn_bwd_kernel (+prep/finish)", "achieved": attn_tflops,
                     "peak": sustained, "peak_kind": f"{which} sustained cuBLAS bf16", "unit": "TFLOP/s",
                     "frac": attn_tflops / sustained, "ms_per_launch": attn_ms_avg, "launches_timed": len(attn_ms),
                     "traffic": ncu_traffic_bytes(name)},
    }
    if world > 1:
        line["allreduce_exposed_ms_per_step"] = exposed_ms
    if breakdown:
        line["breakdown_ms"] = breakdown
    if world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample
        sec, cores, kind = cpu_reference_steps(cfg_kwargs, sample, steps=2, warmup=1)
        line["cpu_baseline"] = {"value": sample / sec, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{sample} sample(s) of the workload's per-sample shape per step, fp32, train "
                                          "mode with the reference's own dropout inside the timed region; the full "
                                          "step (fwd + CE + bwd + AdamW); median of 2 steps after 1 warm-up"}
    if world == 1 and not args.no_eager_baseline:
        del x_dev, y_dev
        torch.cuda.empty_cache()
        try:
            line["gpu_eager_baseline"] = gpu_eager_baseline(cfg_kwargs, dev)
        except Exception as exc:  # the comparator must never take the measurement down with it
            line["gpu_eager_baseline"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

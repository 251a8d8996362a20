"""Generate golden fixtures from the REAL reference (runs only in the build container).

    python tests/golden/make_golden.py

Imports ``temporal_vit.models.model`` from /root/reference (read-only, unmodified), loads fully
randomised parameters, runs forward + CrossEntropy + backward exactly as the
reference training loop does (train.py:167-170, 223-226) and stores inputs / parameters / logits /
per-op intermediates / parameter gradients as small ``.npz`` files next to this script.
Train-mode cases also record the keep-masks the reference's own nn.Dropout / DropPath modules
drew (captured with forward hooks), so the oracle can replay the identical step.

The fixtures are committed; the GPU box has no /root/reference and never runs this script.
"""
from __future__ import annotations

import os
import sys
from dataclasses import asdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from temporal_vit.models import model as ref  # noqa: E402  (the real reference)
from oracle import vit_oracle as O  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def run_case(name, cfg_kwargs, batch, train_mode, seed, class_weight=None, label_smoothing=0.0):
    torch.manual_seed(seed)
    cfg = ref.Temporal3DViTConfig(**cfg_kwargs)
    model = ref.Temporal3DViT(cfg)
    ocfg = O.config_from(cfg)
    params = O.random_params(ocfg, seed=seed + 1)
    missing = model.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.train(train_mode)

    g = torch.Generator().manual_seed(seed + 2)
    x = torch.randn(batch, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g)
    y = torch.randint(0, cfg.n_classes, (batch,), generator=g)

    taps, masks = {}, {}
    hooks = []

    def tap(key):
        def fn(_m, _i, out):
            taps[key] = out.detach().clone()
        return fn

    def drop_mask(key, per_sample=False):
        def fn(m, inp, out):
            if not m.training:
                return
            xin = inp[0].detach()
            if per_sample:   # DropPath: a dropped sample is zero everywhere
                keep = (out.detach() != 0).reshape(out.shape[0], -1).any(dim=1)
            else:            # Dropout: zero output with non-zero input == dropped
                keep = (out.detach() != 0) | (xin == 0)
            masks[key] = keep.to(torch.float32)
        return fn

    hooks.append(model.pos_drop.register_forward_hook(drop_mask("pos_drop")))
    for i, blk in enumerate(model.blocks):
        pre = f"blocks.{i}."
        hooks.append(blk.norm1.register_forward_hook(tap(pre + "norm1")))
        hooks.append(blk.attn.qkv.register_forward_hook(tap(pre + "attn.qkv")))
        hooks.append(blk.attn.register_forward_hook(tap(pre + "attn")))
        hooks.append(blk.mlp.fc1.register_forward_hook(tap(pre + "mlp.fc1")))
        hooks.append(blk.mlp.register_forward_hook(tap(pre + "mlp")))
        hooks.append(blk.register_forward_hook(tap(pre + "out")))
        hooks.append(blk.attn.attn_drop.register_forward_hook(drop_mask(pre + "attn_drop")))
        hooks.append(blk.attn.proj_drop.register_forward_hook(drop_mask(pre + "proj_drop")))
        hooks.append(blk.mlp.drop1.register_forward_hook(drop_mask(pre + "drop1")))
        hooks.append(blk.mlp.drop2.register_forward_hook(drop_mask(pre + "drop2")))
        if isinstance(blk.drop_path1, ref.DropPath):
            hooks.append(blk.drop_path1.register_forward_hook(drop_mask(pre + "drop_path1", True)))
            hooks.append(blk.drop_path2.register_forward_hook(drop_mask(pre + "drop_path2", True)))
    hooks.append(model.head[2].register_forward_hook(drop_mask("head_drop")))

    cw = None if class_weight is None else torch.tensor(class_weight, dtype=torch.float32)
    crit = torch.nn.CrossEntropyLoss(weight=cw, label_smoothing=label_smoothing)
    logits = model(x)
    loss = crit(logits, y)
    loss.backward()
    for h in hooks:
        h.remove()

    out = {"x": _np(x), "y": _np(y), "logits": _np(logits), "loss": _np(loss),
           "train_mode": np.array(int(train_mode)), "label_smoothing": np.array(label_smoothing)}
    if cw is not None:
        out["class_weight"] = _np(cw)
    for k, v in asdict(cfg).items():
        out["cfg." + k] = np.array(v)
    for k, v in params.items():
        out["param." + k] = _np(v)
    for k, v in model.named_parameters():
        out["grad." + k] = _np(v.grad if v.grad is not None else torch.zeros_like(v))
    for k, v in taps.items():
        out["tap." + k] = _np(v)
    for k, v in masks.items():
        out["mask." + k] = _np(v).astype(np.uint8)

    if not train_mode:
        model.eval()
        maps = model.get_attention_maps(x)
        out["attn_map.0"] = _np(maps[0])
        out["attn_map.last"] = _np(maps[-1])

    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: loss={float(loss):.6f} logits[0]={_np(logits)[0]} -> {os.path.getsize(path)/1e6:.2f} MB")


def init_checksums():
    """Reference initialisation under a fixed seed: per-tensor (sum, sum of squares) in float64,
    so the drop-in module can prove `torch.manual_seed(s); Temporal3DViT(cfg)` starts identically."""
    out = {}
    for tag, kw in (("tiny", dict(embed_dim=192, n_heads=3, n_layers=4)),
                    ("small_8x128x256", dict(embed_dim=384, n_heads=6, n_layers=8,
                                             n_trials=8, freq_size=128, time_size=256))):
        torch.manual_seed(1234)
        m = ref.Temporal3DViT(ref.Temporal3DViTConfig(**kw))
        names, sums = [], []
        for k, v in m.state_dict().items():
            names.append(k)
            sums.append([float(v.double().sum()), float((v.double() ** 2).sum()), float(v.numel())])
        out[tag + ".names"] = np.array(names)
        out[tag + ".sums"] = np.array(sums, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "init_checksums.npz"), **out)
    print("init_checksums written")


if __name__ == "__main__":
    torch.set_num_threads(4)
    # G1: eval mode (dropouts inert), two blocks, one head of 64, ragged N=17
    run_case("g1_eval_d64",
             dict(n_trials=4, freq_size=16, time_size=32, embed_dim=64, n_heads=1, n_layers=2,
                  dropout=0.1, attention_dropout=0.1, drop_path=0.1),
             batch=3, train_mode=False, seed=100)
    # G2: train mode with every dropout active, masks recorded, class weights + label smoothing
    run_case("g2_train_d128",
             dict(n_trials=2, freq_size=16, time_size=16, embed_dim=128, n_heads=2, n_layers=2,
                  mlp_ratio=2.0, dropout=0.2, attention_dropout=0.1, drop_path=0.3),
             batch=4, train_mode=True, seed=200, class_weight=[0.7, 1.6], label_smoothing=0.05)
    # G3: train mode, all rates zero (the exact-parity training configuration), no LayerScale
    run_case("g3_train_nodrop_nols",
             dict(n_trials=2, freq_size=16, time_size=32, embed_dim=64, n_heads=1, n_layers=1,
                  dropout=0.0, attention_dropout=0.0, drop_path=0.0, layer_scale_init=0.0),
             batch=2, train_mode=True, seed=300)
    init_checksums()

"""Test infrastructure: rebuild every dropout / DropPath keep mask of one train-mode forward of the drop-in model
from what the model recorded (``model.last_draws``: the Philox seed and the DropPath multipliers) with the numpy
restatement of the library's generator (oracle/dropout_ref.py), in the layout ``oracle.vit_oracle.forward(masks=)``
consumes.  Site ids and element-index rules are the library's contract (include/tvit.h):

    site 1  pos_drop   element = flat index of (B, N, D)          (model.py:224,313 of the reference)
    site 2  head_drop  element = flat index of (B, D)             (:250)
    site 16 (l + 1) + 0  attn_drop  element = ((b H + h) N + q) Np + k,  Np = N rounded up to 16   (:102,113)
    site 16 (l + 1) + 1  proj_drop  element = flat index of (B, N, D)                              (:104,117)
    site 16 (l + 1) + 2  drop1      element = flat index of (B, N, hidden)                         (:138,145)
    site 16 (l + 1) + 3  drop2      element = flat index of (B, N, D)                              (:140,147)
"""
from __future__ import annotations

import dataclasses

import numpy as np
import torch

from oracle import dropout_ref as DR
from oracle import vit_oracle as O


def effective_rate(p: float) -> float:
    """The rate the library realises exactly: round(65536 p) / 65536."""
    return DR.thr16(p) / 65536.0


def oracle_config(cfg) -> "O.OracleConfig":
    """Oracle config whose inverted-dropout scaling 1 / (1 - p) equals the library's 1 / (1 - thr16 / 65536)."""
    oc = O.config_from(cfg)
    return dataclasses.replace(oc, dropout=effective_rate(cfg.dropout),
                               attention_dropout=effective_rate(cfg.attention_dropout))


def build_masks(model, batch: int, device="cpu"):
    cfg = model.config
    draws = model.last_draws
    assert draws is not None, "run a train-mode forward first"
    seed = draws["seed"]
    N, D, H = cfg.n_patches + 1, cfg.embed_dim, cfg.n_heads
    hid = int(D * cfg.mlp_ratio)
    Np = (N + 15) // 16 * 16

    def flat(site, p, shape):
        n = int(np.prod(shape))
        return torch.from_numpy(DR.keep_mask(seed, site, p, 0, n).reshape(shape).astype(np.float32)).to(device)

    m = {}
    if cfg.dropout > 0:
        m["pos_drop"] = flat(1, cfg.dropout, (batch, N, D))
        m["head_drop"] = flat(2, cfg.dropout, (batch, D))
    for i in range(cfg.n_layers):
        pre = f"blocks.{i}."
        base = 16 * (i + 1)
        if cfg.attention_dropout > 0:
            m[pre + "attn_drop"] = flat(base + 0, cfg.attention_dropout, (batch, H, N, Np))[..., :N].contiguous()
        if cfg.dropout > 0:
            m[pre + "proj_drop"] = flat(base + 1, cfg.dropout, (batch, N, D))
            m[pre + "drop1"] = flat(base + 2, cfg.dropout, (batch, N, hid))
            m[pre + "drop2"] = flat(base + 3, cfg.dropout, (batch, N, D))
        s1, s2 = draws["drop_path"][i]
        if s1 is not None:
            m[pre + "drop_path1"] = (s1 > 0).to(torch.float32).to(device)
            m[pre + "drop_path2"] = (s2 > 0).to(torch.float32).to(device)
    return m

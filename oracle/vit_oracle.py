"""CPU/GPU restatement of the reference Temporal 3D ViT forward pass -- TEST INFRASTRUCTURE ONLY.

This file is the *oracle* for the hot path.  It is a from-scratch functional restatement of the
algorithm in the reference's ``temporal_vit/models/model.py`` written with plain tensor algebra
(no ``nn.Module``; every op is spelled out so that per-op intermediates can be compared against the
CUDA kernels).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package ``neural_vit_b200`` never does.

Pinning: the reference has no golden vectors or unit tests for the model (SURVEY.md section 8c:
"parity unpinned" by the reference's own tests).  The oracle is therefore pinned *differentially*:
``tests/golden/make_golden.py`` imports the real reference from ``/root/reference`` in the build
container, runs it on seeded inputs with fully randomised parameters and commits inputs, parameters,
logits, intermediates and parameter gradients under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those files.

Reference lines followed (all in /root/reference/temporal_vit/models/model.py):
  * config / derived sizes ............ :6-47
  * DropPath .......................... :57-71
  * LayerScale ........................ :74-82
  * Attention ......................... :88-119
  * MLP ............................... :122-148
  * TransformerBlock .................. :151-178
  * patch embed / pos embed / forward . :197-202, :276-323
  * get_attention_maps ................ :325-350
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

Tensor = torch.Tensor


@dataclass
class OracleConfig:
    """Mirror of the reference dataclass fields (model.py:6-35); defaults identical."""

    n_trials: int = 8
    freq_size: int = 64
    time_size: int = 128
    patch_trial: int = 2
    patch_freq: int = 8
    patch_time: int = 8
    embed_dim: int = 384
    n_heads: int = 6
    n_layers: int = 8
    mlp_ratio: float = 4.0
    dropout: float = 0.1
    attention_dropout: float = 0.1
    drop_path: float = 0.1
    n_classes: int = 2
    layer_scale_init: float = 1e-4

    @property
    def grid(self) -> Tuple[int, int, int]:
        return (self.n_trials // self.patch_trial,
                self.freq_size // self.patch_freq,
                self.time_size // self.patch_time)

    @property
    def n_patches(self) -> int:  # model.py:37-43
        k, f, t = self.grid
        return k * f * t

    @property
    def patch_dim(self) -> int:  # model.py:45-47
        return self.patch_trial * self.patch_freq * self.patch_time


def config_from(obj) -> OracleConfig:
    """Build an OracleConfig from any object/dict that carries the reference field names."""
    names = OracleConfig.__dataclass_fields__.keys()
    if isinstance(obj, dict):
        return OracleConfig(**{k: obj[k] for k in names if k in obj})
    return OracleConfig(**{k: getattr(obj, k) for k in names if hasattr(obj, k)})


# ----------------------------------------------------------------------------------------------
# elementary ops
# ----------------------------------------------------------------------------------------------

def tubelet_im2col(x: Tensor, cfg: OracleConfig) -> Tensor:
    """(B,K,F,T) -> (B, n, patch_dim).  Conv3d with kernel == stride (model.py:197-202,300-303)
    is a GEMM over non-overlapping tubelets; token i = k'*F'*T' + f'*T' + t' (flatten(2) order),
    patch element order = (dk, df, dt) (Conv3d weight layout (D,1,pk,pf,pt))."""
    B = x.shape[0]
    Kp, Fp, Tp = cfg.grid
    pk, pf, pt = cfg.patch_trial, cfg.patch_freq, cfg.patch_time
    x = x.reshape(B, Kp, pk, Fp, pf, Tp, pt)
    x = x.permute(0, 1, 3, 5, 2, 4, 6)
    return x.reshape(B, Kp * Fp * Tp, pk * pf * pt)


def positional_table(pos_k: Tensor, pos_f: Tensor, pos_t: Tensor) -> Tensor:
    """(n, D) factorised table: pos[k'F'T' + f'T' + t'] = pk[k'] + pf[f'] + pt[t'] (model.py:276-285)."""
    pk, pf, pt = pos_k[0], pos_f[0], pos_t[0]
    tab = pk[:, None, None, :] + pf[None, :, None, :] + pt[None, None, :, :]
    return tab.reshape(-1, tab.shape[-1])


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm(D) semantics: biased variance over the last dim, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (model.py:137, 249)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def dropout_apply(x: Tensor, p: float, mask: Optional[Tensor]) -> Tensor:
    """Inverted dropout with an externally supplied keep mask (1=keep).  ``mask is None`` means
    eval mode / p == 0 (identity)."""
    if mask is None or p == 0.0:
        return x
    return x * mask / (1.0 - p)


def attention(x: Tensor, p: Dict[str, Tensor], prefix: str, n_heads: int,
              attn_p: float = 0.0, attn_mask: Optional[Tensor] = None,
              proj_p: float = 0.0, proj_mask: Optional[Tensor] = None,
              taps: Optional[dict] = None) -> Tensor:
    """model.py:106-119.  qkv rows are ordered [q(all heads); k; v], head h = rows h*hd:(h+1)*hd."""
    B, N, C = x.shape
    hd = C // n_heads
    qkv = x @ p[prefix + "qkv.weight"].T + p[prefix + "qkv.bias"]
    if taps is not None:
        taps[prefix + "qkv"] = qkv
    qkv = qkv.reshape(B, N, 3, n_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    a = torch.softmax(s, dim=-1)
    if taps is not None:
        taps[prefix + "probs"] = a
    a = dropout_apply(a, attn_p, attn_mask)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    if taps is not None:
        taps[prefix + "ctx"] = o
    o = o @ p[prefix + "proj.weight"].T + p[prefix + "proj.bias"]
    return dropout_apply(o, proj_p, proj_mask)


def mlp(x: Tensor, p: Dict[str, Tensor], prefix: str,
        drop_p: float = 0.0, mask1: Optional[Tensor] = None, mask2: Optional[Tensor] = None,
        taps: Optional[dict] = None) -> Tensor:
    """model.py:142-148."""
    h = x @ p[prefix + "fc1.weight"].T + p[prefix + "fc1.bias"]
    if taps is not None:
        taps[prefix + "fc1"] = h
    h = dropout_apply(gelu_erf(h), drop_p, mask1)
    h = h @ p[prefix + "fc2.weight"].T + p[prefix + "fc2.bias"]
    return dropout_apply(h, drop_p, mask2)


def drop_path_rates(cfg: OracleConfig) -> List[float]:
    """model.py:227: linspace(0, drop_path, n_layers) -- block 0 always has rate 0."""
    return [float(v) for v in torch.linspace(0, cfg.drop_path, cfg.n_layers)]


# ----------------------------------------------------------------------------------------------
# whole model
# ----------------------------------------------------------------------------------------------

def embed(x: Tensor, p: Dict[str, Tensor], cfg: OracleConfig) -> Tensor:
    """Patch embed + positional add + CLS prepend (model.py:294-310); no dropout here."""
    if x.dim() == 5:
        x = x[:, 0]
    B = x.shape[0]
    D = cfg.embed_dim
    cols = tubelet_im2col(x, cfg)
    w = p["patch_embed.weight"].reshape(D, cfg.patch_dim)
    tok = cols @ w.T + p["patch_embed.bias"]
    tok = tok + positional_table(p["pos_embed_k"], p["pos_embed_f"], p["pos_embed_t"])
    cls = p["cls_token"].expand(B, 1, D)
    return torch.cat([cls, tok], dim=1)


def forward(x: Tensor, p: Dict[str, Tensor], cfg: OracleConfig,
            masks: Optional[Dict[str, Tensor]] = None,
            taps: Optional[dict] = None) -> Tensor:
    """Full forward (model.py:287-323).

    ``masks`` = None reproduces eval mode (or train mode with every rate 0).  To reproduce a
    train-mode step exactly, pass keep-masks: 'pos_drop' (B,N,D); per block i
    'blocks.i.attn_drop' (B,H,N,N), 'blocks.i.proj_drop' (B,N,D), 'blocks.i.drop1' (B,N,4D),
    'blocks.i.drop2' (B,N,D), 'blocks.i.drop_path1' / 'drop_path2' (B,) and 'head_drop' (B,D).
    """
    masks = masks or {}
    h = embed(x, p, cfg)
    if taps is not None:
        taps["embed"] = h
    h = dropout_apply(h, cfg.dropout, masks.get("pos_drop"))
    rates = drop_path_rates(cfg)
    has_ls = cfg.layer_scale_init > 0
    for i in range(cfg.n_layers):
        pre = f"blocks.{i}."
        y = layer_norm(h, p[pre + "norm1.weight"], p[pre + "norm1.bias"])
        if taps is not None:
            taps[pre + "norm1"] = y
        br = attention(y, p, pre + "attn.", cfg.n_heads,
                       cfg.attention_dropout, masks.get(pre + "attn_drop"),
                       cfg.dropout, masks.get(pre + "proj_drop"), taps)
        if taps is not None:
            taps[pre + "attn"] = br
        if has_ls:
            br = br * p[pre + "ls1.gamma"]
        dp = masks.get(pre + "drop_path1")
        if dp is not None and rates[i] > 0:
            br = br / (1.0 - rates[i]) * dp.reshape(-1, 1, 1)
        h = h + br
        y = layer_norm(h, p[pre + "norm2.weight"], p[pre + "norm2.bias"])
        br = mlp(y, p, pre + "mlp.", cfg.dropout, masks.get(pre + "drop1"), masks.get(pre + "drop2"), taps)
        if taps is not None:
            taps[pre + "mlp"] = br
        if has_ls:
            br = br * p[pre + "ls2.gamma"]
        dp = masks.get(pre + "drop_path2")
        if dp is not None and rates[i] > 0:
            br = br / (1.0 - rates[i]) * dp.reshape(-1, 1, 1)
        h = h + br
        if taps is not None:
            taps[pre + "out"] = h
    h = layer_norm(h, p["norm.weight"], p["norm.bias"])
    c = h[:, 0]
    if taps is not None:
        taps["norm_cls"] = c
    c = c @ p["head.0.weight"].T + p["head.0.bias"]
    c = dropout_apply(gelu_erf(c), cfg.dropout, masks.get("head_drop"))
    return c @ p["head.3.weight"].T + p["head.3.bias"]


def attention_maps(x: Tensor, p: Dict[str, Tensor], cfg: OracleConfig) -> List[Tensor]:
    """model.py:325-350 in eval mode: per-block softmax(q k^T * scale), shape (B,H,N,N)."""
    taps: dict = {}
    forward(x, p, cfg, taps=taps)
    return [taps[f"blocks.{i}.attn.probs"] for i in range(cfg.n_layers)]


def loss_and_grads(x: Tensor, labels: Tensor, p: Dict[str, Tensor], cfg: OracleConfig,
                   class_weight: Optional[Tensor] = None, label_smoothing: float = 0.0,
                   masks: Optional[Dict[str, Tensor]] = None):
    """One training step's fwd + loss + bwd exactly as the reference loop drives it
    (train.py:167-170, 223-226): CrossEntropyLoss(weight, label_smoothing) then autograd."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    logits = forward(x, leaves, cfg, masks=masks)
    loss = torch.nn.functional.cross_entropy(logits, labels, weight=class_weight,
                                             label_smoothing=label_smoothing)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return logits.detach(), loss.detach(), grads


def draw_masks(cfg: OracleConfig, batch: int, generator: Optional[torch.Generator] = None,
               device="cpu") -> Dict[str, Tensor]:
    """Draw the keep-masks one train-mode step of the reference consumes (nn.Dropout at model.py:102,104,
    138,140,224,250 and DropPath at :64-71).  Used by the CPU baseline so that its timed region contains
    the same O(B*H*N*N) Bernoulli work the reference's train mode does."""
    N, D, H = cfg.n_patches + 1, cfg.embed_dim, cfg.n_heads
    hid = int(D * cfg.mlp_ratio)

    def bern(shape, p):
        return (torch.rand(shape, generator=generator, device=device) >= p).to(torch.float32)

    m: Dict[str, Tensor] = {}
    if cfg.dropout > 0:
        m["pos_drop"] = bern((batch, N, D), cfg.dropout)
        m["head_drop"] = bern((batch, D), cfg.dropout)
    rates = drop_path_rates(cfg)
    for i in range(cfg.n_layers):
        pre = f"blocks.{i}."
        if cfg.attention_dropout > 0:
            m[pre + "attn_drop"] = bern((batch, H, N, N), cfg.attention_dropout)
        if cfg.dropout > 0:
            m[pre + "proj_drop"] = bern((batch, N, D), cfg.dropout)
            m[pre + "drop1"] = bern((batch, N, hid), cfg.dropout)
            m[pre + "drop2"] = bern((batch, N, D), cfg.dropout)
        if rates[i] > 0:
            m[pre + "drop_path1"] = bern((batch,), rates[i])
            m[pre + "drop_path2"] = bern((batch,), rates[i])
    return m


def param_shapes(cfg: OracleConfig) -> Dict[str, Tuple[int, ...]]:
    """state_dict layout of the reference, in registration order."""
    D = cfg.embed_dim
    Kp, Fp, Tp = cfg.grid
    hid = int(D * cfg.mlp_ratio)
    s: Dict[str, Tuple[int, ...]] = {
        "pos_embed_k": (1, Kp, D), "pos_embed_f": (1, Fp, D), "pos_embed_t": (1, Tp, D),
        "cls_token": (1, 1, D),
        "patch_embed.weight": (D, 1, cfg.patch_trial, cfg.patch_freq, cfg.patch_time),
        "patch_embed.bias": (D,),
    }
    for i in range(cfg.n_layers):
        pre = f"blocks.{i}."
        s[pre + "norm1.weight"] = (D,)
        s[pre + "norm1.bias"] = (D,)
        s[pre + "attn.qkv.weight"] = (3 * D, D)
        s[pre + "attn.qkv.bias"] = (3 * D,)
        s[pre + "attn.proj.weight"] = (D, D)
        s[pre + "attn.proj.bias"] = (D,)
        if cfg.layer_scale_init > 0:
            s[pre + "ls1.gamma"] = (D,)
        s[pre + "norm2.weight"] = (D,)
        s[pre + "norm2.bias"] = (D,)
        s[pre + "mlp.fc1.weight"] = (hid, D)
        s[pre + "mlp.fc1.bias"] = (hid,)
        s[pre + "mlp.fc2.weight"] = (D, hid)
        s[pre + "mlp.fc2.bias"] = (D,)
        if cfg.layer_scale_init > 0:
            s[pre + "ls2.gamma"] = (D,)
    s["norm.weight"] = (D,)
    s["norm.bias"] = (D,)
    s["head.0.weight"] = (D, D)
    s["head.0.bias"] = (D,)
    s["head.3.weight"] = (cfg.n_classes, D)
    s["head.3.bias"] = (cfg.n_classes,)
    return s


def random_params(cfg: OracleConfig, seed: int = 0, dtype=torch.float32, device="cpu") -> Dict[str, Tensor]:
    """Fully randomised parameters: LayerScale gamma ~ O(1), LN affine, biases and
    positional tables all non-trivial so that errors inside the blocks reach the logits."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if name.endswith("gamma"):
            t = 0.5 + 0.5 * t                     # O(1) layer scale
        elif "norm" in name and name.endswith("weight"):
            t = 1.0 + 0.3 * t
        elif name.endswith("bias"):
            t = 0.1 * t
        elif name.startswith("pos_embed") or name == "cls_token":
            t = 0.5 * t
        elif name == "patch_embed.weight":
            t = t / math.sqrt(cfg.patch_dim)
        else:                                      # Linear weights: ~ 1/sqrt(fan_in)
            t = t / math.sqrt(shape[-1])
        out[name] = t.to(dtype=dtype, device=device)
    return out

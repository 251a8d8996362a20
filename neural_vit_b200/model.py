"""Drop-in Temporal 3D ViT whose forward/backward run on hand-written sm_100a kernels.

Mirrors the reference's public interface for the hot path (temporal_vit/models/model.py):
``Temporal3DViTConfig`` (:6-47), ``CONFIGS`` (:51-55) and ``Temporal3DViT(config)`` with
``forward(x) -> logits`` (:287-323), ``get_attention_maps(x)`` (:325-350), ``.config`` and a
``state_dict()`` with identical keys / shapes / dtypes, so the reference's training loop
(train.py:53-74, 216-235), dataloaders and checkpoint layout work unchanged.

Parameters live in ordinary ``nn`` containers (that is what fixes the checkpoint layout and makes
``torch.manual_seed(s); Temporal3DViT(cfg)`` start from the reference's exact initial weights); none
of those containers' forward methods is ever called.  All arithmetic is done by libtvit_b200.so
through three ``torch.autograd.Function``s (embed, encoder block, head).  One Function per block
keeps gradient production progressive so bucketed all-reduce can overlap with backward.

precision:
  "bf16" (default) -- tcgen05/TMEM/TMA tensor-core kernels, bf16 operands, fp32 accumulate,
                      fp32 residual stream / statistics / parameter gradients.
  "fp32"           -- CUDA-core fp32 verification path (same orchestration, same epilogues);
                      used for the "fp32-path loss within 1e-4" parity gate.
  "bf16_simt"      -- bf16 storage with the CUDA-core engine (debugging aid).
  "bf16_tcgemm"    -- tensor-core GEMMs with the CUDA-core attention (debugging aid).
There is no CPU path: calling forward on a CPU tensor raises.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .gradsink import GradSink


@dataclass
class Temporal3DViTConfig:
    """Same fields, defaults and derived properties as the reference dataclass (model.py:6-47)."""

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
    def n_patches(self) -> int:
        return ((self.n_trials // self.patch_trial) * (self.freq_size // self.patch_freq)
                * (self.time_size // self.patch_time))

    @property
    def patch_dim(self) -> int:
        return self.patch_trial * self.patch_freq * self.patch_time


CONFIGS = {
    "tiny": Temporal3DViTConfig(embed_dim=192, n_heads=3, n_layers=4),
    "small": Temporal3DViTConfig(embed_dim=384, n_heads=6, n_layers=8),
    "base": Temporal3DViTConfig(embed_dim=512, n_heads=8, n_layers=12),
}

# name -> (GEMM engine, attention engine, activation dtype)
_PRECISIONS = {
    "bf16": (L.ENGINE_TCGEN05, L.ENGINE_TCGEN05, L.BF16),
    "fp32": (L.ENGINE_SIMT, L.ENGINE_SIMT, L.F32),
    "bf16_simt": (L.ENGINE_SIMT, L.ENGINE_SIMT, L.BF16),
    "bf16_tcgemm": (L.ENGINE_TCGEN05, L.ENGINE_SIMT, L.BF16),   # debugging aid: isolates the attention kernel
}

# dropout call-site ids (unique per forward; see tvit_dropout in include/tvit.h)
_SITE_POS, _SITE_HEAD = 1, 2


def _site(layer: int, which: int) -> int:
    return 16 * (layer + 1) + which  # which: 0 attn_drop, 1 proj_drop, 2 drop1, 3 drop2


class _Ctx:
    """Per-forward launch context shared by the autograd Functions (plain Python, no tensors)."""

    def __init__(self, engine: int, attn_engine: int, dtype: int, training: bool, seed: int,
                 cfg: Temporal3DViTConfig, sink: Optional[GradSink] = None):
        self.engine = engine
        self.attn_engine = attn_engine
        self.dtype = dtype
        self.training = training
        self.seed = seed
        self.cfg = cfg
        self.sink = sink      # flat gradient buffers the backward kernels accumulate into (None: autograd .grad)

    def drop(self, site: int, p: float):
        if not self.training or p <= 0.0:
            return None
        return (self.seed, site, p)


def _empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device)


def _zeros(shape, device):
    return torch.zeros(shape, dtype=torch.float32, device=device)


class _Shadows:
    """Operand copies of the Linear weights in the activation dtype:
    ``w`` [out,in] for forward, ``wt`` [in,out] (rows of w optionally pre-scaled by the LayerScale
    gamma) for the input-gradient GEMM.

    An entry is valid while its stamp matches: parameter storage pointers, their autograd ``_version`` counters
    and the cache ``generation``.  Every writer this package controls that updates parameters through raw pointers
    bumps ``_version`` (``ops.adamw``, ``FusedAdamW``, the DDP broadcast); ``invalidate()`` is the explicit escape
    hatch for anything else (``Temporal3DViT.invalidate_shadows()``).  ``adopt`` installs shadows somebody else
    already wrote (the fused optimizer re-casts them inside its update kernel)."""

    def __init__(self):
        self._cache: Dict[tuple, tuple] = {}
        self.generation = 0

    def invalidate(self) -> None:
        self.generation += 1

    def _stamp(self, weight, gamma, dtype):
        return (weight.data_ptr(), weight._version, None if gamma is None else (gamma.data_ptr(), gamma._version),
                dtype, str(weight.device), self.generation)

    def adopt(self, key, weight, gamma, dtype: int, w_sh, wt_sh) -> None:
        self._cache[key] = (self._stamp(weight, gamma, dtype), w_sh, wt_sh)

    def peek(self, key):
        hit = self._cache.get(key)
        return None if hit is None else (hit[1], hit[2])

    def get(self, key, weight: torch.Tensor, gamma: Optional[torch.Tensor], dtype: int, need_t: bool):
        w2 = weight.reshape(weight.shape[0], -1)
        stamp = self._stamp(weight, gamma, dtype)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == stamp and (hit[2] is not None or not need_t):
            return hit[1], hit[2]
        R, C = w2.shape
        td = ops.torch_dtype(dtype)
        # reuse the previous buffers when only the contents are stale (keeps pointers stable for the TMA cache)
        w_sh = wt_sh = None
        if hit is not None and hit[0][3] == dtype and hit[0][4] == stamp[4]:
            w_sh, wt_sh = hit[1], hit[2]
        if dtype == L.F32:
            w_sh = w2.detach()
        elif w_sh is None or w_sh.shape != (R, C):
            w_sh = _empty((R, C), td, weight.device)
        if need_t and (wt_sh is None or wt_sh.shape != (C, R)):
            wt_sh = _empty((C, R), td, weight.device)
        if dtype != L.F32 or need_t:
            ops.cast_weight(w2.detach(), R, C, None if gamma is None else gamma.detach(),
                            None if dtype == L.F32 else w_sh, wt_sh if need_t else None, dtype)
        self._cache[key] = (stamp, w_sh, wt_sh)
        return w_sh, wt_sh


def _dst(sinks, i, shape, dev):
    """Destination of parameter-gradient i: its slot in the flat gradient sink, or a fresh zero-filled tensor."""
    if sinks is not None and sinks[i] is not None:
        return sinks[i], True
    return _zeros(shape, dev), False


# ---------------------------------------------------------------------------------------------
# embed: tubelet im2col -> GEMM (+bias +pos +dropout) -> CLS row            (model.py:294-313)
# ---------------------------------------------------------------------------------------------
class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pe_w, pe_b, pos_k, pos_f, pos_t, cls, rt: _Ctx, sh: _Shadows, sinks):
        cfg = rt.cfg
        B = x.shape[0]
        D, P, n = cfg.embed_dim, cfg.patch_dim, cfg.n_patches
        N = n + 1
        Kp, Fp, Tp = (cfg.n_trials // cfg.patch_trial, cfg.freq_size // cfg.patch_freq,
                      cfg.time_size // cfg.patch_time)
        td = ops.torch_dtype(rt.dtype)
        dev = x.device
        cols = _empty((B * n, P), td, dev)
        ops.im2col(x, cols, rt.dtype, B, cfg.n_trials, cfg.freq_size, cfg.time_size, cfg.patch_trial,
                   cfg.patch_freq, cfg.patch_time)
        w_sh, _ = sh.get("patch_embed", pe_w, None, rt.dtype, need_t=False)
        h = _empty((B, N, D), torch.float32, dev)
        drop = rt.drop(_SITE_POS, cfg.dropout)
        ops.gemm(rt.engine, rt.dtype, cols, w_sh, B * n, D, P, epilogue=L.EPI_PATCH_EMBED, out=h, bias=pe_b,
                 pos=(pos_k, pos_f, pos_t), grid3=(Kp, Fp, Tp), drop=drop)
        ops.cls_rows(cls, h, B, N, D, drop)
        ctx.rt, ctx.drop, ctx.grid3 = rt, drop, (Kp, Fp, Tp)
        ctx.pe_shape = pe_w.shape
        ctx.sinks = sinks
        ctx.save_for_backward(cols)
        return h

    @staticmethod
    def backward(ctx, g0):
        (cols,) = ctx.saved_tensors
        rt, cfg = ctx.rt, ctx.rt.cfg
        sk = ctx.sinks[0] if ctx.sinks is not None else None
        Kp, Fp, Tp = ctx.grid3
        g0 = g0.contiguous()
        B, N, D = g0.shape
        n, P = N - 1, cfg.patch_dim
        td = ops.torch_dtype(rt.dtype)
        dev = g0.device
        gtok = _empty((B * n, D), td, dev)
        R = _empty((n, D), torch.float32, dev)
        # parameter order of `sinks`: pe_w, pe_b, pos_k, pos_f, pos_t, cls
        use = sk is not None
        dcls = sk[5] if use else _empty((1, 1, D), torch.float32, dev)
        ops.embed_bwd_prep(g0, B, n, D, ctx.drop, gtok, rt.dtype, R, dcls, accumulate=use)
        dpk = sk[2] if use else _empty((1, Kp, D), torch.float32, dev)
        dpf = sk[3] if use else _empty((1, Fp, D), torch.float32, dev)
        dpt = sk[4] if use else _empty((1, Tp, D), torch.float32, dev)
        dpe_b = sk[1] if use else _empty((D,), torch.float32, dev)
        ops.pos_grad_reduce(R, Kp, Fp, Tp, D, dpk, dpf, dpt, dpe_b, accumulate=use)
        dpe_w = sk[0] if use else _zeros((D, P), dev)
        ops.gemm(rt.engine, rt.dtype, gtok, cols, D, P, B * n, epilogue=L.EPI_ACCUM_F32, out=dpe_w,
                 trans_a=True, trans_b=True)
        if use:
            rt.sink.mark_ready(ctx.sinks[1])
            return (None,) * 10
        return None, dpe_w.reshape(ctx.pe_shape), dpe_b, dpk, dpf, dpt, dcls, None, None, None


# ---------------------------------------------------------------------------------------------
# encoder block (model.py:151-178): x + dp(ls1(attn(ln1 x)));  x + dp(ls2(mlp(ln2 x)))
# ---------------------------------------------------------------------------------------------
# TVIT_ATTN_KEEPBITS=0 makes the backward kernel regenerate the attention dropout masks (A-B timing; same results)
_USE_KEEPBITS = os.environ.get("TVIT_ATTN_KEEPBITS", "1") != "0"
_FUSE_ATTN_CS = os.environ.get("TVIT_ATTN_FUSE_CS", "1") != "0"  # A-B timing of the fused qkv-bias column sums


def _block_forward(h, n1w, n1b, qkvw, qkvb, pw, pb, g1, n2w, n2b, f1w, f1b, f2w, f2b, g2, s1, s2,
                   rt: _Ctx, sh: _Shadows, layer: int, need_grad: bool):
    """All launches of one encoder block's forward.  Returns (h_out, saved intermediates, dropout specs)."""
    cfg = rt.cfg
    B, N, D = h.shape
    M, H = B * N, cfg.n_heads
    hd, hid = D // H, f1w.shape[0]
    E, T = rt.engine, rt.dtype
    td = ops.torch_dtype(T)
    dev = h.device

    qkv_w, qkv_wt = sh.get((layer, "qkv"), qkvw, None, T, need_grad)
    proj_w, proj_wt = sh.get((layer, "proj"), pw, g1, T, need_grad)
    fc1_w, fc1_wt = sh.get((layer, "fc1"), f1w, None, T, need_grad)
    fc2_w, fc2_wt = sh.get((layer, "fc2"), f2w, g2, T, need_grad)

    d_attn = rt.drop(_site(layer, 0), cfg.attention_dropout)
    d_proj = rt.drop(_site(layer, 1), cfg.dropout)
    d_fc1 = rt.drop(_site(layer, 2), cfg.dropout)
    d_fc2 = rt.drop(_site(layer, 3), cfg.dropout)

    y1 = _empty((M, D), td, dev)
    mean1, rstd1 = _empty((M,), torch.float32, dev), _empty((M,), torch.float32, dev)
    ops.ln_fwd(h, D, n1w, n1b, y1, T, mean1, rstd1, M, D)
    qkv = _empty((M, 3 * D), td, dev)
    ops.gemm(E, T, y1, qkv_w, M, 3 * D, D, epilogue=L.EPI_STORE, out=qkv, bias=qkvb)
    ao = _empty((M, D), td, dev)
    lse = _empty((B, H, N), torch.float32, dev)
    # dropout keep-flag cache: written by the forward kernel, read by the backward kernel instead of regenerating the
    # masks (include/tvit.h); only allocated when a backward will follow
    keepbits = ops.attn_keepbits(rt.attn_engine, B, N, H, d_attn, dev) if need_grad and _USE_KEEPBITS else None
    ops.attn_fwd(rt.attn_engine, T, qkv, ao, lse, B, N, H, hd, d_attn, keepbits=keepbits)
    h_mid = torch.empty_like(h)
    ops.gemm(E, T, ao, proj_w, M, D, D, epilogue=L.EPI_RESIDUAL, out=h_mid, bias=pb, resid=h, gamma=g1,
             row_scale=s1, rows_per_group=N, drop=d_proj)

    y2 = _empty((M, D), td, dev)
    mean2, rstd2 = _empty((M,), torch.float32, dev), _empty((M,), torch.float32, dev)
    ops.ln_fwd(h_mid, D, n2w, n2b, y2, T, mean2, rstd2, M, D)
    # d act / d pre-activation (GELU' x dropout multiplier), consumed by GELU_BWD; not produced when no backward follows
    hpre = _empty((M, hid), td, dev) if (need_grad or d_fc1 is not None) else None
    act = _empty((M, hid), td, dev)
    ops.gemm(E, T, y2, fc1_w, M, hid, D, epilogue=L.EPI_BIAS_GELU, out=act, aux=hpre, bias=f1b, drop=d_fc1)
    h_out = torch.empty_like(h)
    ops.gemm(E, T, act, fc2_w, M, D, hid, epilogue=L.EPI_RESIDUAL, out=h_out, bias=f2b, resid=h_mid, gamma=g2,
             row_scale=s2, rows_per_group=N, drop=d_fc2)
    saved = (h, y1, mean1, rstd1, qkv, ao, lse, h_mid, y2, mean2, rstd2, hpre, act,
             n1w, qkvw, pw, pb, g1, n2w, f1w, f2w, f2b, g2, s1, s2, qkv_wt, proj_wt, fc1_wt, fc2_wt, keepbits)
    return h_out, saved, (d_attn, d_proj, d_fc1, d_fc2)


class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, n1w, n1b, qkvw, qkvb, pw, pb, g1, n2w, n2b, f1w, f1b, f2w, f2b, g2, s1, s2,
                rt: _Ctx, sh: _Shadows, layer: int, sinks):
        need_grad = any(ctx.needs_input_grad)
        h_out, saved, drops = _block_forward(h, n1w, n1b, qkvw, qkvb, pw, pb, g1, n2w, n2b, f1w, f1b, f2w, f2b, g2,
                                             s1, s2, rt, sh, layer, need_grad)
        if need_grad:
            ctx.rt, ctx.layer = rt, layer
            ctx.drops = drops
            ctx.has_ls = g1 is not None
            ctx.sinks = sinks
            ctx.save_for_backward(*saved)
        return h_out

    @staticmethod
    def backward(ctx, g_out):
        (h, y1, mean1, rstd1, qkv, ao, lse, h_mid, y2, mean2, rstd2, hpre, act,
         n1w, qkvw, pw, pb, g1, n2w, f1w, f2w, f2b, g2, s1, s2,
         qkv_wt, proj_wt, fc1_wt, fc2_wt, keepbits) = ctx.saved_tensors
        rt, cfg = ctx.rt, ctx.rt.cfg
        sk = ctx.sinks[0] if ctx.sinks is not None else None
        d_attn, d_proj, d_fc1, d_fc2 = ctx.drops
        E, T = rt.engine, rt.dtype
        td = ops.torch_dtype(T)
        g_out = g_out.contiguous()
        B, N, D = g_out.shape
        M, H = B * N, cfg.n_heads
        hd, hid = D // H, f1w.shape[0]
        dev = g_out.device
        f32 = torch.float32
        # index of each parameter in the sink-view tuple
        (I_N1W, I_N1B, I_QKVW, I_QKVB, I_PW, I_PB, I_G1, I_N2W, I_N2B, I_F1W, I_F1B, I_F2W, I_F2B, I_G2) = range(14)
        use = sk is not None

        # ---- MLP branch -------------------------------------------------------------------
        gp2 = _empty((M, D), td, dev)
        if ctx.has_ls:
            cs2 = _zeros((D,), dev)
        else:
            cs2, _ = _dst(sk, I_F2B, (D,), dev)            # without LayerScale colsum(gp) IS the bias gradient
        ops.branch_grad_prep(g_out, M, D, s2, N, d_fc2, gp2, T, cs2)
        dh = _empty((M, hid), td, dev)   # grad wrt fc1 pre-activation
        ops.gemm(E, T, gp2, fc2_wt, M, hid, D, epilogue=L.EPI_GELU_BWD, out=dh, aux=hpre)
        # (the GEMM can fold this column sum into its epilogue -- `colsum=` -- but at this shape the warp reductions
        #  cost the LSU-bound epilogue 0.35 ms against 0.26 ms for the HBM-speed pass below: profiles/r2_kern_v11.log)
        d_f1b, _ = _dst(sk, I_F1B, (hid,), dev)
        ops.colsum(dh, T, M, hid, hid, d_f1b)
        if ctx.has_ls:
            G2 = _zeros((D, hid), dev)
        else:
            G2, _ = _dst(sk, I_F2W, (D, hid), dev)
        ops.gemm(E, T, gp2, act, D, hid, M, epilogue=L.EPI_ACCUM_F32, out=G2, trans_a=True, trans_b=True)
        if ctx.has_ls:
            if use:
                d_f2w, d_g2, d_f2b = sk[I_F2W], sk[I_G2], sk[I_F2B]
            else:
                d_f2w, d_g2, d_f2b = torch.empty_like(f2w), _empty((D,), f32, dev), _empty((D,), f32, dev)
            ops.ls_finalize(G2, f2w, g2, f2b, cs2, d_f2w, d_g2, d_f2b, D, hid, accumulate=use)
        else:
            d_f2w, d_g2, d_f2b = G2, None, cs2
        d_f1w, _ = _dst(sk, I_F1W, (hid, D), dev)
        ops.gemm(E, T, dh, y2, hid, D, M, epilogue=L.EPI_ACCUM_F32, out=d_f1w, trans_a=True, trans_b=True)
        dy2 = _empty((M, D), td, dev)
        ops.gemm(E, T, dh, fc1_wt, M, D, hid, epilogue=L.EPI_STORE, out=dy2)
        del dh
        g_mid = torch.empty_like(g_out)
        d_n2w, _ = _dst(sk, I_N2W, (D,), dev)
        d_n2b, _ = _dst(sk, I_N2B, (D,), dev)
        gp1 = _empty((M, D), td, dev)
        if ctx.has_ls:
            cs1 = _zeros((D,), dev)
        else:
            cs1, _ = _dst(sk, I_PB, (D,), dev)
        ops.ln_bwd(dy2, T, h_mid, D, mean2, rstd2, n2w, g_out, g_mid, D, d_n2w, d_n2b, M, D,
                   gp=gp1, row_scale=s1, rows_per_group=N, drop=d_proj, gp_colsum=cs1)
        del dy2

        # ---- attention branch ---------------------------------------------------------------
        dao = _empty((M, D), td, dev)
        ops.gemm(E, T, gp1, proj_wt, M, D, D, epilogue=L.EPI_STORE, out=dao)
        if ctx.has_ls:
            Gp = _zeros((D, D), dev)
        else:
            Gp, _ = _dst(sk, I_PW, (D, D), dev)
        ops.gemm(E, T, gp1, ao, D, D, M, epilogue=L.EPI_ACCUM_F32, out=Gp, trans_a=True, trans_b=True)
        if ctx.has_ls:
            if use:
                d_pw, d_g1, d_pb = sk[I_PW], sk[I_G1], sk[I_PB]
            else:
                d_pw, d_g1, d_pb = torch.empty_like(pw), _empty((D,), f32, dev), _empty((D,), f32, dev)
            ops.ls_finalize(Gp, pw, g1, pb, cs1, d_pw, d_g1, d_pb, D, D, accumulate=use)
        else:
            d_pw, d_g1, d_pb = Gp, None, cs1
        dqkv = _empty((M, 3 * D), td, dev)
        d_qkvb, _ = _dst(sk, I_QKVB, (3 * D,), dev)
        # colsum(dqkv) (the qkv-bias gradient) is folded into the attention-backward kernels when that is a net win:
        # with dropout it is (whole C2 step 1510-1516 vs 1501-1502 samples/s with the separate 0.2 ms pass,
        # alternating runs on one box), without dropout it is not
        fuse_cs = (d_attn is not None and _FUSE_ATTN_CS) or rt.attn_engine != L.ENGINE_TCGEN05
        ops.attn_bwd(rt.attn_engine, T, qkv, ao, dao, lse, dqkv, B, N, H, hd, d_attn, colsum=d_qkvb if fuse_cs else None,
                     keepbits=keepbits)
        if not fuse_cs:
            ops.colsum(dqkv, T, M, 3 * D, 3 * D, d_qkvb)
        d_qkvw, _ = _dst(sk, I_QKVW, (3 * D, D), dev)
        ops.gemm(E, T, dqkv, y1, 3 * D, D, M, epilogue=L.EPI_ACCUM_F32, out=d_qkvw, trans_a=True, trans_b=True)
        dy1 = _empty((M, D), td, dev)
        ops.gemm(E, T, dqkv, qkv_wt, M, D, 3 * D, epilogue=L.EPI_STORE, out=dy1)
        del dqkv
        g_in = torch.empty_like(g_out)
        d_n1w, _ = _dst(sk, I_N1W, (D,), dev)
        d_n1b, _ = _dst(sk, I_N1B, (D,), dev)
        ops.ln_bwd(dy1, T, h, D, mean1, rstd1, n1w, g_mid, g_in, D, d_n1w, d_n1b, M, D)

        if use:
            rt.sink.mark_ready(ctx.sinks[1])
            return (g_in,) + (None,) * 20
        return (g_in, d_n1w, d_n1b, d_qkvw, d_qkvb, d_pw, d_pb, d_g1, d_n2w, d_n2b, d_f1w, d_f1b, d_f2w, d_f2b,
                d_g2, None, None, None, None, None, None)


# ---------------------------------------------------------------------------------------------
# final norm on the CLS rows + classifier head (model.py:244-252, 320-323); fp32 CUDA-core kernels
# ---------------------------------------------------------------------------------------------
class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, nw, nb, w0, b0, w3, b3, rt: _Ctx, sinks):
        cfg = rt.cfg
        B, N, D = h.shape
        C = w3.shape[0]
        dev = h.device
        f32 = torch.float32
        E, T = L.ENGINE_SIMT, L.F32
        c = _empty((B, D), f32, dev)
        mean, rstd = _empty((B,), f32, dev), _empty((B,), f32, dev)
        ops.ln_fwd(h, N * D, nw, nb, c, T, mean, rstd, B, D)
        zpre, z = _empty((B, D), f32, dev), _empty((B, D), f32, dev)
        drop = rt.drop(_SITE_HEAD, cfg.dropout)
        ops.gemm(E, T, c, w0.detach(), B, D, D, epilogue=L.EPI_BIAS_GELU, out=z, aux=zpre, bias=b0, drop=drop)
        logits = _empty((B, C), f32, dev)
        ops.gemm(E, T, z, w3.detach(), B, C, D, epilogue=L.EPI_STORE, out=logits, bias=b3)
        ctx.drop, ctx.N, ctx.rt, ctx.sinks = drop, N, rt, sinks
        ctx.save_for_backward(h, c, mean, rstd, zpre, z, nw, w0, w3)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        h, c, mean, rstd, zpre, z, nw, w0, w3 = ctx.saved_tensors
        sk = ctx.sinks[0] if ctx.sinks is not None else None
        B, N, D = h.shape
        C = w3.shape[0]
        dev = h.device
        f32 = torch.float32
        E, T = L.ENGINE_SIMT, L.F32
        dlogits = dlogits.contiguous().to(f32)
        # parameter order of `sinks`: nw, nb, w0, b0, w3, b3
        d_w3, _ = _dst(sk, 4, (C, D), dev)
        ops.gemm(E, T, dlogits, z, C, D, B, epilogue=L.EPI_ACCUM_F32, out=d_w3, trans_a=True, trans_b=True)
        d_b3, _ = _dst(sk, 5, (C,), dev)
        ops.colsum(dlogits, T, B, C, C, d_b3)
        dzpre = _empty((B, D), f32, dev)   # (dlogits @ w3) * drop * gelu'(zpre)
        ops.gemm(E, T, dlogits, w3.detach(), B, D, C, epilogue=L.EPI_GELU_BWD, out=dzpre, aux=zpre, trans_b=True)
        d_w0, _ = _dst(sk, 2, (D, D), dev)
        ops.gemm(E, T, dzpre, c, D, D, B, epilogue=L.EPI_ACCUM_F32, out=d_w0, trans_a=True, trans_b=True)
        d_b0, _ = _dst(sk, 3, (D,), dev)
        ops.colsum(dzpre, T, B, D, D, d_b0)
        dc = _empty((B, D), f32, dev)
        ops.gemm(E, T, dzpre, w0.detach(), B, D, D, epilogue=L.EPI_STORE, out=dc, trans_b=True)
        g = _zeros((B, N, D), dev)         # only the CLS rows receive gradient
        d_nw, _ = _dst(sk, 0, (D,), dev)
        d_nb, _ = _dst(sk, 1, (D,), dev)
        ops.ln_bwd(dc, T, h, N * D, mean, rstd, nw, None, g, N * D, d_nw, d_nb, B, D)
        if sk is not None:
            ctx.rt.sink.mark_ready(ctx.sinks[1])
            return (g,) + (None,) * 8
        return g, d_nw, d_nb, d_w0, d_b0, d_w3, d_b3, None, None


# ---------------------------------------------------------------------------------------------
# parameter containers (names fix the checkpoint layout; their forward() is never used)
# ---------------------------------------------------------------------------------------------
class _Gamma(nn.Module):
    def __init__(self, dim: int, init_value: float):
        super().__init__()
        self.gamma = nn.Parameter(init_value * torch.ones(dim))


class _AttnParams(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _MlpParams(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _BlockParams(nn.Module):
    def __init__(self, dim: int, hidden: int, layer_scale_init: float, drop_path: float):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _AttnParams(dim)
        self.ls1 = _Gamma(dim, layer_scale_init) if layer_scale_init > 0 else nn.Identity()
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _MlpParams(dim, hidden)
        self.ls2 = _Gamma(dim, layer_scale_init) if layer_scale_init > 0 else nn.Identity()
        self.drop_path_rate = float(drop_path)

    def tensors(self):
        """Parameters in the argument order of _BlockFn (None where LayerScale is disabled)."""
        g1 = self.ls1.gamma if isinstance(self.ls1, _Gamma) else None
        g2 = self.ls2.gamma if isinstance(self.ls2, _Gamma) else None
        return (self.norm1.weight, self.norm1.bias, self.attn.qkv.weight, self.attn.qkv.bias,
                self.attn.proj.weight, self.attn.proj.bias, g1, self.norm2.weight, self.norm2.bias,
                self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias, g2)


class Temporal3DViT(nn.Module):
    """B200-native Temporal 3D ViT with the reference's constructor and ``forward(x) -> logits``."""

    def __init__(self, config: Temporal3DViTConfig, precision: Optional[str] = None):
        super().__init__()
        self.config = config
        if config.n_trials % config.patch_trial != 0:
            raise ValueError("n_trials must be divisible by patch_trial.")
        if config.freq_size % config.patch_freq != 0:
            raise ValueError("freq_size must be divisible by patch_freq.")
        if config.time_size % config.patch_time != 0:
            raise ValueError("time_size must be divisible by patch_time.")
        if config.embed_dim % config.n_heads != 0:
            raise ValueError("embed_dim must be divisible by n_heads.")
        precision = precision or os.environ.get("TVIT_PRECISION", "bf16")
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {precision!r}")
        self.precision = precision

        D = config.embed_dim
        self.patch_embed = nn.Conv3d(1, D, kernel_size=(config.patch_trial, config.patch_freq, config.patch_time),
                                     stride=(config.patch_trial, config.patch_freq, config.patch_time))
        self.n_patches_k = config.n_trials // config.patch_trial
        self.n_patches_f = config.freq_size // config.patch_freq
        self.n_patches_t = config.time_size // config.patch_time
        self.pos_embed_k = nn.Parameter(torch.zeros(1, self.n_patches_k, D))
        self.pos_embed_f = nn.Parameter(torch.zeros(1, self.n_patches_f, D))
        self.pos_embed_t = nn.Parameter(torch.zeros(1, self.n_patches_t, D))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))

        rates = [float(v) for v in torch.linspace(0, config.drop_path, config.n_layers)]
        hidden = int(D * config.mlp_ratio)
        self.blocks = nn.ModuleList(
            [_BlockParams(D, hidden, config.layer_scale_init, rates[i]) for i in range(config.n_layers)])
        self.norm = nn.LayerNorm(D)
        self.head = nn.Sequential(nn.Linear(D, D), nn.GELU(), nn.Dropout(config.dropout),
                                  nn.Linear(D, config.n_classes))
        self._reset_parameters()
        self._shadows = _Shadows()
        self._grad_sink: Optional[GradSink] = None
        # what the last train-mode forward drew: {"seed": int, "drop_path": [(s1, s2) per block]} -- lets the parity
        # tests rebuild every dropout / DropPath mask with oracle/dropout_ref.py
        self.last_draws: Optional[dict] = None

    # same initialisation recipe and RNG consumption order as the reference (model.py:257-274)
    def _reset_parameters(self) -> None:
        for p in (self.pos_embed_k, self.pos_embed_f, self.pos_embed_t, self.cls_token):
            nn.init.trunc_normal_(p, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    # ------------------------------------------------------------------------------------------
    # hooks for the optimizer / data-parallel layers of this package
    # ------------------------------------------------------------------------------------------
    def invalidate_shadows(self) -> None:
        """Force the bf16 operand copies of the weights to be rebuilt at the next forward.  Needed only after the
        parameters were modified by something that neither bumps their autograd version counter nor is part of this
        package (the fused optimizer and the DDP broadcast do it themselves)."""
        self._shadows.invalidate()

    def attach_grad_sink(self, sink: Optional[GradSink] = None, bucket_mb: float = 32.0) -> GradSink:
        """Make the backward kernels accumulate parameter gradients straight into flat buckets (gradsink.py)."""
        if self._grad_sink is None:
            self._grad_sink = sink if sink is not None else GradSink(list(self.parameters()), bucket_mb)
        return self._grad_sink

    def detach_grad_sink(self) -> None:
        self._grad_sink = None

    def _apply(self, fn, *args, **kwargs):
        if self._grad_sink is not None:
            raise RuntimeError("move / cast the model before creating FusedAdamW or BucketedAllReduce: its parameters "
                               "and gradients live in flat buffers now")
        return super()._apply(fn, *args, **kwargs)

    # ------------------------------------------------------------------------------------------
    def _runtime(self, x: torch.Tensor) -> _Ctx:
        if not x.is_cuda:
            raise RuntimeError("Temporal3DViT (B200 build) runs on CUDA tensors only; there is no CPU fallback. "
                               "Move the model and inputs to a B200 (sm_100a) device.")
        L.require_device(x.device.index if x.device.index is not None else torch.cuda.current_device())
        if self.patch_embed.weight.device != x.device:
            raise RuntimeError("model parameters and input are on different devices")
        engine, attn_engine, dtype = _PRECISIONS[self.precision]
        cfg = self.config
        if engine == L.ENGINE_TCGEN05:
            if attn_engine == L.ENGINE_TCGEN05 and cfg.embed_dim // cfg.n_heads != 64:
                raise ValueError("the tcgen05 attention kernel supports head_dim == 64 only "
                                 f"(embed_dim={cfg.embed_dim}, n_heads={cfg.n_heads})")
            if cfg.patch_dim % 8 != 0 or cfg.embed_dim % 8 != 0 or int(cfg.embed_dim * cfg.mlp_ratio) % 8 != 0:
                raise ValueError("the tcgen05 path needs patch_dim, embed_dim and the MLP width to be multiples of 8")
        if cfg.embed_dim % 4 != 0:
            raise ValueError("embed_dim must be a multiple of 4")
        seed = 0
        if self.training and (cfg.dropout > 0.0 or cfg.attention_dropout > 0.0):
            hi, lo = torch.randint(0, 2 ** 31 - 1, (2,)).tolist()   # CPU generator: honours torch.manual_seed
            seed = (hi << 31) | lo
        sink = self._grad_sink if (self.training and torch.is_grad_enabled()) else None
        if sink is not None:
            sink.begin_pass()
        return _Ctx(engine, attn_engine, dtype, self.training, seed, cfg, sink)

    def _check_input(self, x: torch.Tensor) -> torch.Tensor:
        cfg = self.config
        if x.dim() == 5:
            if x.shape[1] != 1:
                raise ValueError(f"expected a single input channel, got shape {tuple(x.shape)}")
            x = x[:, 0]
        if x.dim() != 4 or tuple(x.shape[1:]) != (cfg.n_trials, cfg.freq_size, cfg.time_size):
            raise ValueError(f"expected input (B,{cfg.n_trials},{cfg.freq_size},{cfg.time_size}) or with a channel "
                             f"dim of 1, got {tuple(x.shape)}")
        return x.to(torch.float32).contiguous()

    def _drop_path_scales(self, B: int, device):
        """Per-sample DropPath multipliers floor(keep + U[0,1)) / keep (model.py:64-71) for both branches of every
        block, drawn with ONE torch.rand per forward; entries are None where the rate is 0 / in eval mode."""
        Lyr = self.config.n_layers
        rates = [blk.drop_path_rate for blk in self.blocks]
        if not self.training or not any(r > 0.0 for r in rates):
            return [(None, None)] * Lyr
        keep = getattr(self, "_dp_keep", None)
        if keep is None or keep.device != device:
            keep = torch.tensor([1.0 - r for r in rates for _ in (0, 1)], dtype=torch.float32, device=device)[:, None]
            self._dp_keep = keep
        scales = torch.floor(keep + torch.rand(2 * Lyr, B, device=device, dtype=torch.float32)) / keep
        return [(scales[2 * i], scales[2 * i + 1]) if rates[i] > 0.0 else (None, None) for i in range(Lyr)]

    def _sinks_for(self, rt: _Ctx, params):
        """(gradient-sink views, the Parameter objects themselves) in the Function's argument order, or None."""
        if rt.sink is None:
            return None
        return tuple(None if p is None else rt.sink.view(p) for p in params), tuple(p for p in params if p is not None)

    def _embed(self, x: torch.Tensor, rt: _Ctx) -> torch.Tensor:
        params = (self.patch_embed.weight, self.patch_embed.bias, self.pos_embed_k, self.pos_embed_f,
                  self.pos_embed_t, self.cls_token)
        return _EmbedFn.apply(x, *params, rt, self._shadows, self._sinks_for(rt, params))

    def _block(self, i: int, h: torch.Tensor, rt: _Ctx, dp) -> torch.Tensor:
        params = self.blocks[i].tensors()
        return _BlockFn.apply(h, *params, dp[0], dp[1], rt, self._shadows, i, self._sinks_for(rt, params))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, K, F, T) or (B, 1, K, F, T) fp32 on a B200 -> logits (B, n_classes) fp32."""
        if not x.is_cuda:
            self._runtime(x)   # raises: no CPU fallback
        with torch.cuda.device(x.device), torch.autocast(device_type="cuda", enabled=False):
            rt = self._runtime(x)
            x = self._check_input(x)
            dps = self._drop_path_scales(x.shape[0], x.device)
            if self.training:
                self.last_draws = {"seed": rt.seed, "drop_path": dps}
            h = self._embed(x, rt)
            for i in range(self.config.n_layers):
                h = self._block(i, h, rt, dps[i])
            params = (self.norm.weight, self.norm.bias, self.head[0].weight, self.head[0].bias,
                      self.head[3].weight, self.head[3].bias)
            return _HeadFn.apply(h, *params, rt, self._sinks_for(rt, params))

    def get_attention_maps(self, x: torch.Tensor) -> List[torch.Tensor]:
        """Per-block softmax(q k^T / sqrt(hd)) of shape (B, H, N, N), as the reference (model.py:325-350).

        Tensor-core path: each block runs its normal forward launches (qkv GEMM, flash attention -> row
        log-sum-exp), then the probabilities are materialised tile by tile by the tcgen05 GEMM with the
        SOFTMAX_PROBS epilogue, P = exp(scale * Q K^T - lse), one launch per (sample, head).  The fp32
        verification precision uses the CUDA-core row-softmax kernel."""
        if not x.is_cuda:
            self._runtime(x)
        maps: List[torch.Tensor] = []
        with torch.cuda.device(x.device), torch.autocast(device_type="cuda", enabled=False), torch.no_grad():
            rt = self._runtime(x)
            x = self._check_input(x)
            cfg = self.config
            D, H = cfg.embed_dim, cfg.n_heads
            hd = D // H
            dps = self._drop_path_scales(x.shape[0], x.device)
            h = self._embed(x, rt)
            B, N, _ = h.shape
            for i, blk in enumerate(self.blocks):
                h_out, saved, _ = _block_forward(h, *blk.tensors(), dps[i][0], dps[i][1], rt, self._shadows, i, False)
                qkv, lse = saved[4], saved[6]
                probs = torch.empty((B, H, N, N), dtype=torch.float32, device=h.device)
                if rt.attn_engine == L.ENGINE_TCGEN05:
                    lse_flat = lse.reshape(-1)
                    for b in range(B):
                        for hh in range(H):
                            ops.gemm(rt.engine, rt.dtype, qkv, qkv, N, N, hd, epilogue=L.EPI_SOFTMAX_PROBS, out=probs,
                                     lda=3 * D, ldb=3 * D, ldo=N, row_scale=lse_flat, alpha=hd ** -0.5,
                                     a_offset=b * N * 3 * D + hh * hd, b_offset=b * N * 3 * D + D + hh * hd,
                                     out_offset=(b * H + hh) * N * N, row_scale_offset=(b * H + hh) * N)
                else:
                    ops.attn_probs(rt.dtype, qkv, probs, B, N, H, hd)
                maps.append(probs)
                h = h_out
        return maps

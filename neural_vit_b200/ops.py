"""Thin tensor-level wrappers over the C ABI (include/tvit.h).

Every function takes CUDA tensors allocated by PyTorch, passes raw device pointers and the current
CUDA stream to libtvit_b200.so, and raises RuntimeError on any non-zero return code.  No arithmetic
happens in Python.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib as L

DropSpec = Optional[Tuple[int, int, float]]  # (seed, site, p)

# --- instrumentation used by bench.py (off by default) -------------------------------------------------
# LAUNCHES counts the CUDA kernels this library launched (per C call: see _KERNELS_PER_CALL);
# when TIMED is a dict {name: [(start_event, end_event), ...]} the named ops are bracketed by CUDA events
# on the launching stream so a kernel's duration can be measured inside a timed training step.
LAUNCHES = {"count": 0}
TIMED = None
_KERNELS_PER_CALL = {"gemm": 1, "attn_fwd": 1, "attn_bwd": 3, "attn_bwd_simt": 2, "attn_probs": 1, "im2col": 1,
                     "ln_fwd": 1, "ln_bwd": 1, "branch_grad_prep": 1, "colsum": 1, "cast_weight": 1,
                     "ls_finalize": 1, "cls_rows": 1, "embed_bwd_prep": 1, "pos_grad_reduce": 1, "adamw": 1,
                     "shadow_t_multi": 1, "ce_loss": 1}


def _count(name: str) -> None:
    LAUNCHES["count"] += _KERNELS_PER_CALL[name]


class _timed:
    def __init__(self, name):
        self.name = name
        self.on = TIMED is not None and (name in TIMED or "detail" in TIMED)
        if self.on and name not in TIMED:
            TIMED[name] = []

    def __enter__(self):
        if self.on:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()

    def __exit__(self, *exc):
        if self.on:
            self.e.record()
            TIMED[self.name].append((self.s, self.e))
        return False

_TORCH_DTYPE = {L.F32: torch.float32, L.BF16: torch.bfloat16}


def torch_dtype(code: int) -> torch.dtype:
    return _TORCH_DTYPE[code]


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _drop(spec: DropSpec) -> L.Dropout:
    if spec is None:
        return L.Dropout(0, 0, 0.0)
    seed, site, p = spec
    return L.Dropout(int(seed), int(site), float(p))


def _drop_ptr(spec: DropSpec):
    return ctypes.byref(_drop(spec))


def _req(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the B200 kernels have no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")


def gemm(engine: int, dtype: int, a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, *, epilogue: int,
         out: torch.Tensor, lda: Optional[int] = None, ldb: Optional[int] = None, ldo: Optional[int] = None,
         trans_a: bool = False, trans_b: bool = False, bias=None, aux=None, resid=None, gamma=None, row_scale=None,
         rows_per_group: int = 0, drop: DropSpec = None, pos=None, grid3=None, split_k: int = 0,
         alpha: float = 1.0, a_offset: int = 0, b_offset: int = 0, out_offset: int = 0,
         row_scale_offset: int = 0, colsum=None) -> None:
    """C[M,N] = op(A) op(B)^T with a fused epilogue (see tvit_gemm in include/tvit.h)."""
    td = torch_dtype(dtype)
    _req(a, td, "gemm A")
    _req(b, td, "gemm B")
    args = L.GemmArgs()
    args.engine, args.dtype = engine, dtype
    args.trans_a, args.trans_b = int(trans_a), int(trans_b)
    args.M, args.N, args.K = M, N, K
    # *_offset: element offsets into the operand tensors (sub-matrix views, e.g. one head of the packed qkv)
    args.A, args.lda = a.data_ptr() + a_offset * a.element_size(), (lda if lda is not None else (M if trans_a else K))
    args.B, args.ldb = b.data_ptr() + b_offset * b.element_size(), (ldb if ldb is not None else (N if trans_b else K))
    args.epilogue = epilogue
    args.out, args.ldo = out.data_ptr() + out_offset * out.element_size(), (ldo if ldo is not None else N)
    args.bias = _ptr(bias)
    args.aux, args.ldaux = _ptr(aux), N
    args.resid, args.ldres = _ptr(resid), N
    args.gamma = _ptr(gamma)
    args.row_scale = None if row_scale is None else row_scale.data_ptr() + 4 * row_scale_offset
    args.rows_per_group = rows_per_group
    args.drop = _drop(drop)
    if pos is not None:
        args.pos_k, args.pos_f, args.pos_t = (p.data_ptr() for p in pos)
        args.Kp, args.Fp, args.Tp = grid3
    args.split_k = split_k
    args.alpha = alpha
    args.colsum = _ptr(colsum)
    _count("gemm")
    with _timed(("gemm_e%d%s" % (epilogue, "_tn" if trans_a else "")) if TIMED is None or "detail" not in TIMED
                else "gemm_e%d%s_N%d_K%d" % (epilogue, "_tn" if trans_a else "", N, K)):
        L.check(L.load().tvit_gemm(ctypes.byref(args), _stream()), "tvit_gemm")


def attn_keepbits(engine, B, N, H, drop: DropSpec, device) -> Optional[torch.Tensor]:
    """Buffer for the dropout keep-flag cache that attn_fwd fills and attn_bwd reads (include/tvit.h), or None when
    the engine / call does not use one."""
    if drop is None or drop[2] <= 0.0:
        return None
    nbytes = int(L.load().tvit_attn_keepbits_bytes(engine, B, N, H))
    return torch.empty(nbytes, dtype=torch.uint8, device=device) if nbytes else None


def attn_fwd(engine, dtype, qkv, out, lse, B, N, H, hd, drop: DropSpec = None, keepbits=None) -> None:
    _count("attn_fwd")
    with _timed("attn_fwd"):
        L.check(L.load().tvit_attn_fwd(engine, dtype, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, hd,
                                       _drop_ptr(drop), _ptr(keepbits), _stream()), "tvit_attn_fwd")


# grow-only scratch buffers, one per (device, stream): every use is stream-ordered on the stream it is keyed by, so
# consecutive backward launches reuse the same block without a trip through the allocator
_WORKSPACES = {}


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    key = (device.index, _stream())
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def attn_bwd(engine, dtype, qkv, out, dout, lse, dqkv, B, N, H, hd, drop: DropSpec = None, colsum=None,
             keepbits=None) -> None:
    lib = L.load()
    nbytes = int(lib.tvit_attn_bwd_workspace_bytes(engine, dtype, B, N, H, hd))
    ws = workspace(nbytes, qkv.device)
    _count("attn_bwd" if engine == L.ENGINE_TCGEN05 else "attn_bwd_simt")
    with _timed("attn_bwd"):
        L.check(lib.tvit_attn_bwd(engine, dtype, qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                  dqkv.data_ptr(), ws.data_ptr(), nbytes, B, N, H, hd, _drop_ptr(drop), _ptr(colsum),
                                  _ptr(keepbits), _stream()),
                "tvit_attn_bwd")


def attn_probs(dtype, qkv, probs, B, N, H, hd) -> None:
    _count("attn_probs")
    L.check(L.load().tvit_attn_probs(dtype, qkv.data_ptr(), probs.data_ptr(), B, N, H, hd, _stream()),
            "tvit_attn_probs")


def im2col(x, cols, dtype, B, K, F, T, pk, pf, pt) -> None:
    _req(x, torch.float32, "im2col x")
    _count("im2col")
    L.check(L.load().tvit_im2col(x.data_ptr(), cols.data_ptr(), dtype, B, K, F, T, pk, pf, pt, _stream()),
            "tvit_im2col")


def ln_fwd(x, x_row_stride, weight, bias, y, dtype, mean, rstd, rows, D, eps=1e-5) -> None:
    _count("ln_fwd")
    L.check(L.load().tvit_ln_fwd(x.data_ptr(), x_row_stride, weight.data_ptr(), bias.data_ptr(), y.data_ptr(), dtype,
                                 _ptr(mean), _ptr(rstd), rows, D, eps, _stream()), "tvit_ln_fwd")


def ln_bwd(dy, dtype, x, x_row_stride, mean, rstd, weight, g_res, dx, dx_row_stride, dweight, dbias, rows, D, *,
           gp=None, row_scale=None, rows_per_group=0, drop: DropSpec = None, gp_colsum=None) -> None:
    _count("ln_bwd")
    L.check(L.load().tvit_ln_bwd(dy.data_ptr(), dtype, x.data_ptr(), x_row_stride, mean.data_ptr(), rstd.data_ptr(),
                                 weight.data_ptr(), _ptr(g_res), dx.data_ptr(), dx_row_stride, _ptr(dweight),
                                 _ptr(dbias), _ptr(gp), _ptr(row_scale), rows_per_group, _drop_ptr(drop),
                                 _ptr(gp_colsum), rows, D, _stream()), "tvit_ln_bwd")


def branch_grad_prep(g, rows, D, row_scale, rows_per_group, drop: DropSpec, gp, dtype, colsum) -> None:
    _count("branch_grad_prep")
    L.check(L.load().tvit_branch_grad_prep(g.data_ptr(), rows, D, _ptr(row_scale), rows_per_group, _drop_ptr(drop),
                                           gp.data_ptr(), dtype, _ptr(colsum), _stream()), "tvit_branch_grad_prep")


def colsum(x, dtype, rows, C, ld, out) -> None:
    _count("colsum")
    L.check(L.load().tvit_colsum(x.data_ptr(), dtype, rows, C, ld, out.data_ptr(), _stream()), "tvit_colsum")


def cast_weight(w, R, C, row_scale, out, out_t, dtype) -> None:
    _count("cast_weight")
    L.check(L.load().tvit_cast_weight(w.data_ptr(), R, C, _ptr(row_scale), _ptr(out), _ptr(out_t), dtype, _stream()),
            "tvit_cast_weight")


def ls_finalize(G, W, gamma, bias, cs, dW, dgamma, dbias, R, C, accumulate: bool = False) -> None:
    _count("ls_finalize")
    L.check(L.load().tvit_ls_finalize(G.data_ptr(), _ptr(W), _ptr(gamma), _ptr(bias), cs.data_ptr(), dW.data_ptr(),
                                      _ptr(dgamma), _ptr(dbias), R, C, int(accumulate), _stream()), "tvit_ls_finalize")


def cls_rows(cls, h, B, N, D, drop: DropSpec) -> None:
    _count("cls_rows")
    L.check(L.load().tvit_cls_rows(cls.data_ptr(), h.data_ptr(), B, N, D, _drop_ptr(drop), _stream()),
            "tvit_cls_rows")


def embed_bwd_prep(g0, B, n, D, drop: DropSpec, gtok, dtype, R, dcls, accumulate: bool = False) -> None:
    _count("embed_bwd_prep")
    L.check(L.load().tvit_embed_bwd_prep(g0.data_ptr(), B, n, D, _drop_ptr(drop), gtok.data_ptr(), dtype,
                                         R.data_ptr(), dcls.data_ptr(), int(accumulate), _stream()),
            "tvit_embed_bwd_prep")


def pos_grad_reduce(R, Kp, Fp, Tp, D, dpk, dpf, dpt, dbias, accumulate: bool = False) -> None:
    _count("pos_grad_reduce")
    L.check(L.load().tvit_pos_grad_reduce(R.data_ptr(), Kp, Fp, Tp, D, dpk.data_ptr(), dpf.data_ptr(),
                                          dpt.data_ptr(), dbias.data_ptr(), int(accumulate), _stream()),
            "tvit_pos_grad_reduce")


def adamw(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, shadow=None) -> None:
    """Fused AdamW over flat fp32 tensors (p is updated through its raw pointer: the autograd version counter is
    bumped here so that every cache keyed on ``_version`` -- the operand shadows -- sees the change)."""
    _req(p, torch.float32, "adamw p")
    _req(g, torch.float32, "adamw g")
    if shadow is not None:
        _req(shadow, torch.bfloat16, "adamw shadow")
    _count("adamw")
    L.check(L.load().tvit_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(shadow), p.numel(), lr,
                                beta1, beta2, eps, weight_decay, step, grad_scale, _stream()), "tvit_adamw")
    torch.autograd.graph.increment_version(p)


def shadow_t_multi(descs_device: torch.Tensor, count: int, total_tiles: int, dtype: int) -> None:
    _count("shadow_t_multi")
    L.check(L.load().tvit_shadow_t_multi(descs_device.data_ptr(), count, total_tiles, dtype, _stream()),
            "tvit_shadow_t_multi")


def ce_loss(logits, labels, class_weight, label_smoothing, loss, dlogits, metric_acc=None, prob_out=None,
            label_out=None) -> None:
    _req(logits, torch.float32, "ce_loss logits")
    _req(labels, torch.int64, "ce_loss labels")
    B, C = logits.shape
    _count("ce_loss")
    L.check(L.load().tvit_ce_loss(logits.data_ptr(), labels.data_ptr(), _ptr(class_weight), float(label_smoothing), B,
                                  C, _ptr(loss), _ptr(dlogits), _ptr(metric_acc), _ptr(prob_out), _ptr(label_out),
                                  _stream()), "tvit_ce_loss")

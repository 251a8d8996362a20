"""Kernel-level parity (B200 only): every C-ABI op against a plain fp32/fp64 PyTorch restatement of
the same arithmetic on the same seeded inputs.  Tolerances: fp32 engine ~1e-5 relative; bf16 tensor
core engine: inputs are pre-rounded to bf16 so only accumulation order and the bf16 output rounding
differ (<= 1e-2 relative Frobenius, bf16 eps = 3.9e-3)."""
import math

import pytest
import torch

from neural_vit_b200 import _lib as L
from neural_vit_b200 import ops
from oracle import vit_oracle as O
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

ENGINES = [(L.ENGINE_SIMT, L.F32, "simt_f32"), (L.ENGINE_SIMT, L.BF16, "simt_bf16"),
           (L.ENGINE_TCGEN05, L.BF16, "tc_bf16")]


def _tol(dtype):
    return 2e-5 if dtype == L.F32 else 1e-2


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = (torch.randn(shape, generator=g) * scale).to(DEV)
    return t.to(ops.torch_dtype(dtype)).contiguous()


GEMM_SHAPES = [
    (128, 128, 64), (256, 384, 384), (300, 192, 128), (2049 * 2, 1152, 384), (515, 64, 64),
    (1000, 1536, 384), (777, 384, 1536), (4098, 128, 128), (130, 256, 512), (64, 576, 192),
]


@pytest.mark.parametrize("engine,dtype,tag", ENGINES, ids=[e[2] for e in ENGINES])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_store_bias(engine, dtype, tag, M, N, K):
    a, b = _rand((M, K), dtype, 1), _rand((N, K), dtype, 2, 1 / math.sqrt(K))
    bias = _rand((N,), L.F32, 3)
    out = torch.empty((M, N), dtype=ops.torch_dtype(dtype), device=DEV)
    ops.gemm(engine, dtype, a, b, M, N, K, epilogue=L.EPI_STORE, out=out, bias=bias)
    ref = a.double() @ b.double().T + bias.double()
    assert rel_err(out, ref) < _tol(dtype)


@pytest.mark.parametrize("engine,dtype,tag", ENGINES, ids=[e[2] for e in ENGINES])
@pytest.mark.parametrize("M,N,K", [(384, 1536, 2049 * 2), (1152, 384, 5000), (128, 128, 64), (384, 384, 777),
                                    (64, 128, 300), (192, 768, 1025)])
def test_gemm_wgrad_accum(engine, dtype, tag, M, N, K):
    """C[M,N] += A^T B with A stored [K,M], B stored [K,N] (weight-gradient layout, split along K)."""
    a, b = _rand((K, M), dtype, 4), _rand((K, N), dtype, 5, 1 / math.sqrt(K))
    out = torch.zeros((M, N), dtype=torch.float32, device=DEV)
    ops.gemm(engine, dtype, a, b, M, N, K, epilogue=L.EPI_ACCUM_F32, out=out, trans_a=True, trans_b=True)
    ref = a.double().T @ b.double()
    assert rel_err(out, ref) < (2e-5 if dtype == L.F32 else 1e-4)
    # accumulate semantics: a second call adds
    ops.gemm(engine, dtype, a, b, M, N, K, epilogue=L.EPI_ACCUM_F32, out=out, trans_a=True, trans_b=True)
    assert rel_err(out, 2 * ref) < (2e-5 if dtype == L.F32 else 1e-4)


@pytest.mark.parametrize("engine,dtype,tag", ENGINES, ids=[e[2] for e in ENGINES])
def test_gemm_bias_gelu(engine, dtype, tag):
    M, N, K = 700, 1536, 384
    a, b = _rand((M, K), dtype, 6), _rand((N, K), dtype, 7, 1 / math.sqrt(K))
    bias = _rand((N,), L.F32, 8)
    td = ops.torch_dtype(dtype)
    out, aux = torch.empty((M, N), dtype=td, device=DEV), torch.empty((M, N), dtype=td, device=DEV)
    ops.gemm(engine, dtype, a, b, M, N, K, epilogue=L.EPI_BIAS_GELU, out=out, aux=aux, bias=bias)
    h = (a.double() @ b.double().T + bias.double()).requires_grad_(True)
    act = O.gelu_erf(h)
    act.sum().backward()
    assert rel_err(out, act) < _tol(dtype)
    # aux = d out / d h (GELU' here; times the dropout multiplier when dropout is on): what GELU_BWD multiplies by
    assert rel_err(aux, h.grad) < _tol(dtype)
    # with dropout both outputs carry the SAME mask and the 1/(1-p) factor
    drop = (77, 4, 0.25)
    out_d, aux_d = torch.empty_like(out), torch.empty_like(aux)
    ops.gemm(engine, dtype, a, b, M, N, K, epilogue=L.EPI_BIAS_GELU, out=out_d, aux=aux_d, bias=bias, drop=drop)
    keep = aux_d != 0                       # gelu' vanishes nowhere on these inputs (|h| < 6)
    assert abs(keep.float().mean().item() - 0.75) < 5e-3
    assert torch.equal(out_d.float()[~keep], torch.zeros_like(out_d.float()[~keep]))
    assert rel_err(out_d.float()[keep], out.float()[keep] / 0.75) < 2 * _tol(dtype) + 1e-3
    assert rel_err(aux_d.float()[keep], aux.float()[keep] / 0.75) < 2 * _tol(dtype) + 1e-3
    # an inference forward passes no aux buffer: the same `out` bits, nothing else written
    out_n = torch.empty_like(out)
    ops.gemm(engine, dtype, a, b, M, N, K, epilogue=L.EPI_BIAS_GELU, out=out_n, bias=bias)
    assert torch.equal(out_n, out)
    # ... also on the weight-stationary tcgen05 path (enough m-tiles for every SM) and with a ragged last m-tile
    if engine == L.ENGINE_TCGEN05:
        M2 = 128 * 2 * 148 + 77
        a2 = _rand((M2, K), dtype, 60)
        o1, o2, x1 = (torch.empty((M2, N), dtype=td, device=DEV) for _ in range(3))
        ops.gemm(engine, dtype, a2, b, M2, N, K, epilogue=L.EPI_BIAS_GELU, out=o1, aux=x1, bias=bias)
        ops.gemm(engine, dtype, a2, b, M2, N, K, epilogue=L.EPI_BIAS_GELU, out=o2, bias=bias)
        assert torch.equal(o1, o2)


@pytest.mark.parametrize("engine,dtype,tag", ENGINES, ids=[e[2] for e in ENGINES])
def test_gemm_residual_layerscale_droppath(engine, dtype, tag):
    Bsz, Ntok, N, K = 3, 171, 384, 384
    M = Bsz * Ntok
    a, b = _rand((M, K), dtype, 9), _rand((N, K), dtype, 10, 1 / math.sqrt(K))
    bias, gamma = _rand((N,), L.F32, 11), _rand((N,), L.F32, 12)
    resid = _rand((M, N), L.F32, 13)
    rs = torch.tensor([0.0, 1.25, 1.25], device=DEV)
    out = torch.empty((M, N), dtype=torch.float32, device=DEV)
    ops.gemm(engine, dtype, a, b, M, N, K, epilogue=L.EPI_RESIDUAL, out=out, bias=bias, resid=resid, gamma=gamma,
             row_scale=rs, rows_per_group=Ntok)
    z = a.double() @ b.double().T + bias.double()
    ref = resid.double() + rs.double().repeat_interleave(Ntok)[:, None] * gamma.double() * z
    assert rel_err(out, ref) < (2e-5 if dtype == L.F32 else 2e-4)
    # without gamma / row_scale
    ops.gemm(engine, dtype, a, b, M, N, K, epilogue=L.EPI_RESIDUAL, out=out, bias=bias, resid=resid)
    assert rel_err(out, resid.double() + z) < (2e-5 if dtype == L.F32 else 2e-4)


@pytest.mark.parametrize("engine,dtype,tag", ENGINES, ids=[e[2] for e in ENGINES])
def test_gemm_gelu_bwd(engine, dtype, tag):
    M, N, K = 520, 1536, 384
    a, b = _rand((M, K), dtype, 14), _rand((N, K), dtype, 15, 1 / math.sqrt(K))
    daux = _rand((M, N), dtype, 16)      # d act / d pre-activation as written by the BIAS_GELU epilogue
    out = torch.empty((M, N), dtype=ops.torch_dtype(dtype), device=DEV)
    ops.gemm(engine, dtype, a, b, M, N, K, epilogue=L.EPI_GELU_BWD, out=out, aux=daux)
    ref = (a.double() @ b.double().T) * daux.double()
    assert rel_err(out, ref) < _tol(dtype)
    # optional fused column sums of the output (the bias gradient of the preceding Linear); accumulate semantics
    for (M2, N2) in ((M, N), (40000, 1536), (300, 200)):
        a2, d2 = _rand((M2, K), dtype, 17), _rand((M2, N2), dtype, 18)
        b2 = _rand((N2, K), dtype, 19, 1 / math.sqrt(K))
        out2 = torch.empty((M2, N2), dtype=ops.torch_dtype(dtype), device=DEV)
        cs = torch.ones(N2, device=DEV)
        ops.gemm(engine, dtype, a2, b2, M2, N2, K, epilogue=L.EPI_GELU_BWD, out=out2, aux=d2, colsum=cs)
        ref2 = (a2.double() @ b2.double().T) * d2.double()
        assert rel_err(out2, ref2) < _tol(dtype)
        assert rel_err(cs - 1.0, ref2.sum(0)) < (1e-4 if dtype == L.F32 else 2e-3), (M2, N2)


@pytest.mark.parametrize("engine,dtype,tag", ENGINES, ids=[e[2] for e in ENGINES])
def test_gemm_patch_embed(engine, dtype, tag):
    cfg = O.OracleConfig(n_trials=4, freq_size=32, time_size=64, embed_dim=192, n_heads=3)
    Bsz = 3
    Kp, Fp, Tp = cfg.grid
    n, P, D = cfg.n_patches, cfg.patch_dim, cfg.embed_dim
    x = _rand((Bsz, 4, 32, 64), L.F32, 17)
    p = {k: v.to(DEV) for k, v in O.random_params(cfg, seed=5).items()}
    td = ops.torch_dtype(dtype)
    cols = torch.empty((Bsz * n, P), dtype=td, device=DEV)
    ops.im2col(x, cols, dtype, Bsz, 4, 32, 64, 2, 8, 8)
    assert rel_err(cols, O.tubelet_im2col(x, cfg).reshape(Bsz * n, P)) < (1e-7 if dtype == L.F32 else 4e-3)
    w = p["patch_embed.weight"].reshape(D, P).to(td).contiguous()
    h = torch.full((Bsz, n + 1, D), float("nan"), dtype=torch.float32, device=DEV)
    ops.gemm(engine, dtype, cols, w, Bsz * n, D, P, epilogue=L.EPI_PATCH_EMBED, out=h,
             bias=p["patch_embed.bias"], pos=(p["pos_embed_k"], p["pos_embed_f"], p["pos_embed_t"]),
             grid3=(Kp, Fp, Tp))
    ops.cls_rows(p["cls_token"], h, Bsz, n + 1, D, None)
    pr = dict(p)
    pr["patch_embed.weight"] = w.float().reshape(p["patch_embed.weight"].shape)
    xin = O.tubelet_im2col(x, cfg).to(td).float()          # same operand rounding as the kernel
    tok = xin.double() @ w.double().T + p["patch_embed.bias"].double()
    tok = tok + O.positional_table(p["pos_embed_k"], p["pos_embed_f"], p["pos_embed_t"]).double()
    ref = torch.cat([p["cls_token"].double().expand(Bsz, 1, D), tok], dim=1)
    assert rel_err(h, ref) < (2e-5 if dtype == L.F32 else 1e-4)


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
@pytest.mark.parametrize("rows,D", [(1000, 384), (77, 64), (513, 768), (300, 192), (64, 1024), (50, 128)])
def test_layernorm_fwd_bwd(dtype, rows, D):
    x = _rand((rows, D), L.F32, 20, 2.0) + 0.5
    w, b = _rand((D,), L.F32, 21) * 0.3 + 1.0, _rand((D,), L.F32, 22)
    td = ops.torch_dtype(dtype)
    y = torch.empty((rows, D), dtype=td, device=DEV)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    ops.ln_fwd(x, D, w, b, y, dtype, mean, rstd, rows, D)
    xd = x.double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd, (D,), wd, bd, 1e-5)
    assert rel_err(y, ref) < (1e-6 if dtype == L.F32 else 4e-3)
    assert rel_err(mean, xd.mean(-1)) < 1e-6
    dy = _rand((rows, D), dtype, 23)
    gres = _rand((rows, D), L.F32, 24)
    ref.backward(dy.double())
    dx = torch.empty((rows, D), device=DEV)
    dw, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    rs = torch.rand(5, device=DEV) + 0.5
    rpg = (rows + 4) // 5
    gp = torch.empty((rows, D), dtype=td, device=DEV)
    gpcs = torch.zeros(D, device=DEV)
    ops.ln_bwd(dy, dtype, x, D, mean, rstd, w, gres, dx, D, dw, db, rows, D, gp=gp, row_scale=rs,
               rows_per_group=rpg, gp_colsum=gpcs)
    dx_ref = gres.double() + xd.grad
    assert rel_err(dx, dx_ref) < 2e-5
    assert rel_err(dw, wd.grad) < 2e-5 and rel_err(db, bd.grad) < 2e-5
    gp_ref = dx_ref * rs.double().repeat_interleave(rpg)[:rows, None]
    assert rel_err(gp, gp_ref) < (2e-5 if dtype == L.F32 else 4e-3)
    assert rel_err(gpcs, gp_ref.sum(0)) < 2e-5


def test_layernorm_strided_cls_rows():
    Bsz, Ntok, D = 5, 33, 128
    h = _rand((Bsz, Ntok, D), L.F32, 30)
    w, b = _rand((D,), L.F32, 31), _rand((D,), L.F32, 32)
    y = torch.empty((Bsz, D), device=DEV)
    mean, rstd = torch.empty(Bsz, device=DEV), torch.empty(Bsz, device=DEV)
    ops.ln_fwd(h, Ntok * D, w, b, y, L.F32, mean, rstd, Bsz, D)
    ref = torch.nn.functional.layer_norm(h[:, 0].double(), (D,), w.double(), b.double(), 1e-5)
    assert rel_err(y, ref) < 1e-6


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
def test_branch_grad_prep_colsum_cast(dtype):
    rows, D, rpg = 700, 384, 100
    g = _rand((rows, D), L.F32, 40)
    rs = torch.rand(7, device=DEV)
    td = ops.torch_dtype(dtype)
    gp = torch.empty((rows, D), dtype=td, device=DEV)
    cs = torch.zeros(D, device=DEV)
    ops.branch_grad_prep(g, rows, D, rs, rpg, None, gp, dtype, cs)
    ref = g.double() * rs.double().repeat_interleave(rpg)[:, None]
    assert rel_err(gp, ref) < (1e-6 if dtype == L.F32 else 4e-3)
    assert rel_err(cs, ref.sum(0)) < 2e-5
    out = torch.zeros(D, device=DEV)
    ops.colsum(gp, dtype, rows, D, D, out)
    assert rel_err(out, gp.double().sum(0)) < 2e-5
    # ragged tiny C (logit bias gradient)
    t = _rand((37, 2), L.F32, 41)
    o2 = torch.zeros(2, device=DEV)
    ops.colsum(t, L.F32, 37, 2, 2, o2)
    assert rel_err(o2, t.double().sum(0)) < 1e-6
    # weight shadows
    w = _rand((300, 130), L.F32, 42)
    sc = _rand((300,), L.F32, 43)
    wc = torch.empty((300, 130), dtype=td, device=DEV)
    wt = torch.empty((130, 300), dtype=td, device=DEV)
    ops.cast_weight(w, 300, 130, sc, wc, wt, dtype)
    assert rel_err(wc, w) < (1e-7 if dtype == L.F32 else 4e-3)
    assert rel_err(wt, (w * sc[:, None]).T) < (1e-7 if dtype == L.F32 else 4e-3)


def test_ls_finalize_matches_autograd():
    R, C, M = 96, 160, 50
    a = _rand((M, C), L.F32, 50).double()
    W = _rand((R, C), L.F32, 51).double().requires_grad_(True)
    b = _rand((R,), L.F32, 52).double().requires_grad_(True)
    gam = _rand((R,), L.F32, 53).double().requires_grad_(True)
    gp = _rand((M, R), L.F32, 54).double()
    ((a @ W.T + b) * gam * gp).sum().backward()
    G = (gp.T @ a).float().contiguous()
    cs = gp.sum(0).float().contiguous()
    dW, dg, db = (torch.empty((R, C), device=DEV), torch.empty(R, device=DEV), torch.empty(R, device=DEV))
    ops.ls_finalize(G, W.detach().float().contiguous(), gam.detach().float().contiguous(),
                    b.detach().float().contiguous(), cs, dW, dg, db, R, C)
    assert rel_err(dW, W.grad) < 1e-5 and rel_err(dg, gam.grad) < 1e-5 and rel_err(db, b.grad) < 1e-5


def _attn_ref(qkv, Bsz, N, H, hd, mask=None, p=0.0):
    D = H * hd
    t = qkv.double().reshape(Bsz, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    a = torch.softmax((q @ k.transpose(-2, -1)) * hd ** -0.5, dim=-1)
    lse = torch.logsumexp((q @ k.transpose(-2, -1)) * hd ** -0.5, dim=-1)
    if mask is not None:
        a = a * mask / (1 - p)
    return (a @ v).transpose(1, 2).reshape(Bsz * N, D), lse


ATTN_ENGINES = ENGINES
ATTN_SHAPES = [(2, 17, 1, 64), (2, 257, 3, 64), (1, 513, 2, 64), (3, 130, 6, 64), (1, 2049, 1, 64)]


@pytest.mark.parametrize("engine,dtype,tag", ATTN_ENGINES, ids=[e[2] for e in ATTN_ENGINES])
@pytest.mark.parametrize("Bsz,N,H,hd", ATTN_SHAPES)
def test_attention_fwd_bwd(engine, dtype, tag, Bsz, N, H, hd):
    D = H * hd
    td = ops.torch_dtype(dtype)
    qkv = _rand((Bsz * N, 3 * D), dtype, 60, 1.0)
    out = torch.empty((Bsz * N, D), dtype=td, device=DEV)
    lse = torch.empty((Bsz, H, N), device=DEV)
    ops.attn_fwd(engine, dtype, qkv, out, lse, Bsz, N, H, hd)
    qd = qkv.double().requires_grad_(True)
    ref, lse_ref = _attn_ref(qd, Bsz, N, H, hd)
    assert rel_err(out, ref) < (2e-5 if dtype == L.F32 else 1e-2)
    assert rel_err(lse, lse_ref) < 2e-5 if dtype == L.F32 else rel_err(lse, lse_ref) < 1e-3
    dout = _rand((Bsz * N, D), dtype, 61)
    ref.backward(dout.double())
    dqkv = torch.empty_like(qkv)
    cs = torch.zeros(3 * D, device=DEV)
    ops.attn_bwd(engine, dtype, qkv, out, dout, lse, dqkv, Bsz, N, H, hd, colsum=cs)
    tol = 5e-5 if dtype == L.F32 else 2e-2
    g = qd.grad.reshape(Bsz * N, 3, D)
    got = dqkv.reshape(Bsz * N, 3, D)
    for i, nm in enumerate("qkv"):
        assert rel_err(got[:, i], g[:, i]) < tol, f"d{nm}"
    # fused qkv-bias gradient = column sums of dqkv; the K third is zero in exact arithmetic (softmax shift invariance)
    want = qd.grad.sum(0)
    for i in (0, 2):
        assert rel_err(cs[i * D:(i + 1) * D], want[i * D:(i + 1) * D]) < (1e-4 if dtype == L.F32 else 1e-2), "qkv"[i]
    assert cs[D:2 * D].abs().max().item() < 1e-2 * cs.abs().max().item() + 1e-6


@pytest.mark.parametrize("Bsz,N,H", [(2, 300, 2), (1, 2049, 1), (3, 130, 6), (2, 17, 1)])
def test_attention_dropout_mask_is_the_same_function_in_every_kernel(Bsz, N, H):
    """The dropout mask is a pure function of (seed, site, b, h, q, k).  The tcgen05 forward and backward kernels
    index it by tile / key half / 16-key group, the SIMT kernels element by element: with the same descriptor all
    four must realise the same mask, so outputs and gradients agree to bf16 accuracy.  A mask that was shifted by
    one group in one kernel would pass every statistical test and fail this one."""
    hd, p = 64, 0.3
    D = H * hd
    drop = (424242, 9, p)
    qkv = _rand((Bsz * N, 3 * D), L.BF16, 80, 0.7)
    dout = _rand((Bsz * N, D), L.BF16, 81)
    res = {}
    for eng in (L.ENGINE_SIMT, L.ENGINE_TCGEN05):
        out = torch.empty((Bsz * N, D), dtype=torch.bfloat16, device=DEV)
        lse = torch.empty((Bsz, H, N), device=DEV)
        ops.attn_fwd(eng, L.BF16, qkv, out, lse, Bsz, N, H, hd, drop)
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(eng, L.BF16, qkv, out, dout, lse, dqkv, Bsz, N, H, hd, drop)
        res[eng] = (out.float(), lse.clone(), dqkv.float())
    a, b = res[L.ENGINE_SIMT], res[L.ENGINE_TCGEN05]
    assert rel_err(b[0], a[0]) < 1e-2          # a different mask at p = 0.3 would give O(0.5)
    assert rel_err(b[1], a[1]) < 1e-3
    assert rel_err(b[2], a[2]) < 2e-2
    # and the mask really is there: without dropout the outputs differ grossly
    out0 = torch.empty((Bsz * N, D), dtype=torch.bfloat16, device=DEV)
    ops.attn_fwd(L.ENGINE_TCGEN05, L.BF16, qkv, out0, torch.empty((Bsz, H, N), device=DEV), Bsz, N, H, hd)
    assert rel_err(b[0], out0.float()) > 0.1


@pytest.mark.parametrize("variant", [3, 12], ids=["transposed", "whole_tile_two_issuers"])
@pytest.mark.parametrize("Bsz,N,H", [(1, 2049, 2), (2, 300, 3), (2, 129, 1)])
def test_attention_backward_variants_agree(variant, Bsz, N, H):
    """The tcgen05 backward kernel exists in three formulations (tvit_attn_bwd_variant, include/tvit.h): key-half
    pipelined (default), transposed (P^T / dS^T as TMEM operands; its dropout masks go through an in-register 16 x 16
    byte transpose) and whole-tile S / dP with two MMA-issuing warps.  All must produce the default's gradients -- with
    dropout that means the identical mask -- and the SIMT engine's, for full and ragged tiles."""
    hd = 64
    D = H * hd
    qkv = _rand((Bsz * N, 3 * D), L.BF16, 90, 0.7)
    dout = _rand((Bsz * N, D), L.BF16, 91)
    lib = L.load()
    for drop in (None, (99, 5, 0.3)):
        out = torch.empty((Bsz * N, D), dtype=torch.bfloat16, device=DEV)
        lse = torch.empty((Bsz, H, N), device=DEV)
        ops.attn_fwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, lse, Bsz, N, H, hd, drop)
        got = {}
        prev = lib.tvit_attn_bwd_variant(-1)
        try:
            for v in (0, variant):
                lib.tvit_attn_bwd_variant(v)
                dqkv = torch.full_like(qkv, float("nan"))
                cs = torch.zeros(3 * D, device=DEV)
                ops.attn_bwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, dout, lse, dqkv, Bsz, N, H, hd, drop, colsum=cs)
                got[v] = (dqkv.float(), cs.clone())
        finally:
            lib.tvit_attn_bwd_variant(prev)
        ref = torch.empty_like(qkv)
        ops.attn_bwd(L.ENGINE_SIMT, L.BF16, qkv, out, dout, lse, ref, Bsz, N, H, hd, drop)
        assert torch.isfinite(got[variant][0]).all()
        # same operands, same masks, same bf16 roundings of P and dS: the variants differ only in summation order
        assert rel_err(got[variant][0], got[0][0]) < 2e-3, drop
        assert rel_err(got[variant][0], ref.float()) < 2e-2, drop
        assert rel_err(got[variant][1], got[0][1]) < 2e-3, drop


@pytest.mark.parametrize("Bsz,N,H,p", [(1, 2049, 2, 0.1), (2, 300, 3, 0.3), (2, 129, 1, 0.1), (1, 257, 2, 0.75),
                                       (2, 128, 1, 0.5), (1, 77, 2, 0.1)])
def test_attention_keepbits_cache_reproduces_the_generated_masks(Bsz, N, H, p):
    """tvit_attn_fwd can record the dropout keep flags it applied (include/tvit.h: tvit_attn_keepbits_bytes) and
    tvit_attn_bwd can read them instead of running the generator again.  The cache is the same function of
    (seed, site, element), so the backward result must be BIT-identical with and without it -- for full, ragged and
    128 m + 1 token counts (where the tail key's flag takes the scalar path), for p above 1/2, and with the forward
    output unchanged by the recording."""
    hd = 64
    D = H * hd
    qkv = _rand((Bsz * N, 3 * D), L.BF16, 92, 0.7)
    dout = _rand((Bsz * N, D), L.BF16, 93)
    drop = (1234, 7, p)
    E = L.ENGINE_TCGEN05
    out0 = torch.empty((Bsz * N, D), dtype=torch.bfloat16, device=DEV)
    out1 = torch.empty_like(out0)
    lse0 = torch.empty((Bsz, H, N), device=DEV)
    lse1 = torch.empty_like(lse0)
    kb = ops.attn_keepbits(E, Bsz, N, H, drop, torch.device(DEV))
    assert kb is not None and kb.numel() == Bsz * H * ((N + 127) // 128) ** 2 * 128 * 8 * 2
    # the cache sits between two guard regions that the forward kernel must leave untouched; stale garbage inside
    # it must not leak into valid elements
    guard = 4096
    arena = torch.full((kb.numel() + 2 * guard,), 0xA5, dtype=torch.uint8, device=DEV)
    kb = arena[guard:guard + kb.numel()]
    ops.attn_fwd(E, L.BF16, qkv, out0, lse0, Bsz, N, H, hd, drop)
    ops.attn_fwd(E, L.BF16, qkv, out1, lse1, Bsz, N, H, hd, drop, keepbits=kb)
    assert torch.equal(out0, out1) and torch.equal(lse0, lse1)
    assert (arena[:guard] == 0xA5).all() and (arena[guard + kb.numel():] == 0xA5).all()
    d0 = torch.full_like(qkv, float("nan"))
    d1 = torch.full_like(qkv, float("nan"))
    ops.attn_bwd(E, L.BF16, qkv, out0, dout, lse0, d0, Bsz, N, H, hd, drop)
    ops.attn_bwd(E, L.BF16, qkv, out0, dout, lse0, d1, Bsz, N, H, hd, drop, keepbits=kb)
    assert torch.isfinite(d1.float()).all()
    # dK / dV are reduced inside one CTA in a fixed order: bit-identical.  dQ is accumulated with fp32 atomics across
    # the key-tile CTAs, whose order differs from launch to launch: equal to rounding.
    assert torch.equal(d0[:, D:], d1[:, D:])
    assert rel_err(d1[:, :D].float(), d0[:, :D].float()) < 1e-3
    # the keep rate recorded in the cache (valid query rows / keys only) is 1 - p
    nt = (N + 127) // 128
    words = kb.view(torch.int16).view(Bsz * H, nt, nt, 128, 8).to(torch.int32)
    bits = torch.stack([(words >> s) & 1 for s in range(16)], dim=-1)  # [..., 8 groups, 16 bits]
    # element 2t -> bit t, element 2t+1 -> bit 8+t: reorder to key order
    order = [i // 2 + (8 if i % 2 else 0) for i in range(16)]
    bits = bits[..., order].reshape(Bsz * H, nt, nt, 128, 128)
    full = bits.permute(0, 1, 3, 2, 4).reshape(Bsz * H, nt * 128, nt * 128)[:, :N, :N].float()
    assert abs(full.mean().item() - (1 - p)) < 4.0 * (p * (1 - p) / full.numel()) ** 0.5 + 1e-3
    # ... and it is the oracle's mask (numpy restatement of the generator)
    from oracle import dropout_ref
    Np = (N + 15) // 16 * 16
    ref0 = dropout_ref.keep_mask(drop[0], drop[1], p, 0, N * Np).reshape(N, Np)[:, :N]  # (b, h) = (0, 0): q * Np + k
    assert (torch.from_numpy(ref0.astype("float32")) == full[0].cpu()).all()


@pytest.mark.parametrize("engine,dtype,tag", ATTN_ENGINES, ids=[e[2] for e in ATTN_ENGINES])
def test_attention_dropout_consistency(engine, dtype, tag):
    """With dropout the mask cannot match torch's Philox stream; check keep-rate, 1/(1-p) scaling and
    that forward and backward use the SAME mask (gradient check against the recovered mask)."""
    Bsz, N, H, hd, p = 2, 129, 2, 64, 0.25
    D = H * hd
    td = ops.torch_dtype(dtype)
    qkv = _rand((Bsz * N, 3 * D), dtype, 70, 0.5)
    # make V = identity-like probe so the dropped probabilities can be read back: use v = one-hot rows
    out = torch.empty((Bsz * N, D), dtype=td, device=DEV)
    lse = torch.empty((Bsz, H, N), device=DEV)
    drop = (1234567, 16, p)
    ops.attn_fwd(engine, dtype, qkv, out, lse, Bsz, N, H, hd, drop)
    out2 = torch.empty_like(out)
    ops.attn_fwd(engine, dtype, qkv, out2, lse, Bsz, N, H, hd, drop)
    assert torch.equal(out, out2)                       # deterministic in (seed, site)
    out3 = torch.empty_like(out)
    ops.attn_fwd(engine, dtype, qkv, out3, lse, Bsz, N, H, hd, (7654321, 16, p))
    assert not torch.equal(out, out3)
    # expectation over many seeds approaches the no-dropout output
    acc = torch.zeros_like(out, dtype=torch.float32)
    nrep = 48
    for s in range(nrep):
        ops.attn_fwd(engine, dtype, qkv, out3, lse, Bsz, N, H, hd, (1000 + s, 16, p))
        acc += out3.float()
    ops.attn_fwd(engine, dtype, qkv, out2, lse, Bsz, N, H, hd)
    assert rel_err(acc / nrep, out2.float()) < 0.15


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
def test_elementwise_dropout_statistics_and_replay(dtype):
    rows, D, p = 2000, 384, 0.2
    g = torch.ones((rows, D), device=DEV)
    td = ops.torch_dtype(dtype)
    gp = torch.empty((rows, D), dtype=td, device=DEV)
    cs = torch.zeros(D, device=DEV)
    ops.branch_grad_prep(g, rows, D, None, 1, (99, 5, p), gp, dtype, cs)
    keep = (gp.float() != 0)
    rate = keep.float().mean().item()
    assert abs(rate - (1 - p)) < 5e-3
    vals = gp.float()[keep]
    assert torch.allclose(vals, torch.full_like(vals, 1 / (1 - p)), rtol=4e-3)
    # the RESIDUAL epilogue with the same (seed, site) draws the same mask
    a = torch.zeros((rows, 64), dtype=td, device=DEV)
    b = torch.zeros((D, 64), dtype=td, device=DEV)
    bias = torch.ones(D, device=DEV)
    out = torch.empty((rows, D), device=DEV)
    resid = torch.zeros((rows, D), device=DEV)
    eng = L.ENGINE_SIMT if dtype == L.F32 else L.ENGINE_TCGEN05
    ops.gemm(eng, dtype, a, b, rows, D, 64, epilogue=L.EPI_RESIDUAL, out=out, bias=bias, resid=resid,
             drop=(99, 5, p))
    assert torch.equal(out != 0, keep)


@pytest.mark.parametrize("p", [0.1, 0.123, 0.5, 1.0 / 256.0, 0.9])
def test_dropout_rate_is_exact_despite_8bit_lanes(p):
    """16 elements share one Philox call (8 random bits each); the per-group threshold dither must make the drop
    probability equal p to ~2^-16, not p rounded to 1/256: over 2^24 elements the binomial std of the keep rate is
    below 1.3e-4, while p = 0.123 rounded to 31/256 or 32/256 would be off by 1.9e-3 / 2e-3."""
    rows, D = 1 << 14, 1 << 10
    g = torch.ones((rows, D), device=DEV)
    gp = torch.empty((rows, D), dtype=torch.float32, device=DEV)
    cs = torch.zeros(D, device=DEV)
    ops.branch_grad_prep(g, rows, D, None, 1, (20240607, 3, p), gp, L.F32, cs)
    keep = gp != 0
    rate = keep.double().mean().item()
    assert abs(rate - (1 - p)) < 6e-4, rate
    vals = gp[keep]
    assert torch.allclose(vals, torch.full_like(vals, 1 / (1 - p)), rtol=1e-4)
    # no structure along rows or columns: per-column and per-row keep rates are binomial around 1 - p
    col, row = keep.double().mean(0), keep.double().mean(1)
    sd_c, sd_r = math.sqrt(p * (1 - p) / rows), math.sqrt(p * (1 - p) / D)
    assert (col - (1 - p)).abs().max().item() < 6 * sd_c + 1e-4
    assert (row - (1 - p)).abs().max().item() < 6 * sd_r + 1e-4
    # neighbouring elements inside a 16-element group are (almost) uncorrelated
    k = keep.double()
    c = ((k[:, :-1] - (1 - p)) * (k[:, 1:] - (1 - p))).mean().item() / (p * (1 - p))
    assert abs(c) < 2e-3, c


@pytest.mark.parametrize("seed,site,p", [(20240607, 3, 0.123), ((7 << 40) + 12345, 17, 0.1), (1, 0, 0.5)])
def test_dropout_mask_matches_numpy_oracle(seed, site, p):
    """Integer work, so the bar is bit-exactness: the masks realised by the elementwise kernel, by the tcgen05 GEMM
    epilogues (fp32 RESIDUAL path and packed-bf16 BIAS_GELU path) and by both attention engines equal
    oracle/dropout_ref.keep_mask element for element."""
    from oracle import dropout_ref as R
    rows, D = 300, 384
    ref = torch.from_numpy(R.keep_mask(seed, site, p, 0, rows * D).reshape(rows, D)).to(DEV)
    g = torch.ones((rows, D), device=DEV)
    gp = torch.empty((rows, D), dtype=torch.float32, device=DEV)
    cs = torch.zeros(D, device=DEV)
    ops.branch_grad_prep(g, rows, D, None, 1, (seed, site, p), gp, L.F32, cs)
    assert torch.equal(gp != 0, ref)
    kept = gp[ref]
    assert torch.allclose(kept, torch.full_like(kept, R.inv_keep(p)), rtol=1e-6)
    # LayerNorm backward's fused branch-gradient output (quad-shared Philox words): dy = 0 -> dx = g_res = 1
    for dt in (L.F32, L.BF16):
        td = ops.torch_dtype(dt)
        dy0 = torch.zeros((rows, D), dtype=td, device=DEV)
        xln = torch.randn((rows, D), device=DEV)
        mean, rstd = torch.zeros(rows, device=DEV), torch.ones(rows, device=DEV)
        dx = torch.empty((rows, D), device=DEV)
        gpl = torch.empty((rows, D), dtype=td, device=DEV)
        ops.ln_bwd(dy0, dt, xln, D, mean, rstd, torch.ones(D, device=DEV), g, dx, D, torch.zeros(D, device=DEV),
                   torch.zeros(D, device=DEV), rows, D, gp=gpl, row_scale=None, rows_per_group=1,
                   drop=(seed, site, p), gp_colsum=torch.zeros(D, device=DEV))
        assert torch.equal(gpl != 0, ref), dt
    # tcgen05 GEMM epilogues: zero operands, bias 1 -> the output is the mask times a constant
    a = torch.zeros((rows, 64), dtype=torch.bfloat16, device=DEV)
    b = torch.zeros((D, 64), dtype=torch.bfloat16, device=DEV)
    bias = torch.ones(D, device=DEV)
    out = torch.empty((rows, D), device=DEV)
    ops.gemm(L.ENGINE_TCGEN05, L.BF16, a, b, rows, D, 64, epilogue=L.EPI_RESIDUAL, out=out, bias=bias,
             resid=torch.zeros((rows, D), device=DEV), drop=(seed, site, p))
    assert torch.equal(out != 0, ref)
    out16 = torch.empty((rows, D), dtype=torch.bfloat16, device=DEV)
    aux16 = torch.empty_like(out16)
    ops.gemm(L.ENGINE_TCGEN05, L.BF16, a, b, rows, D, 64, epilogue=L.EPI_BIAS_GELU, out=out16, aux=aux16, bias=bias,
             drop=(seed, site, p))
    assert torch.equal(out16 != 0, ref) and torch.equal(aux16 != 0, ref)
    # attention: V = identity-like probe is not needed -- with q = k = 0 every probability is 1/N, so
    # out[b, q, h, :] = inv_keep / N * sum_k keep(b,h,q,k) v[k]; use v[k] = one-hot(k mod 64) and compare the counts
    Bsz, N, H, hd = 2, 150, 2, 64
    Dm = H * hd
    qkv = torch.zeros((Bsz * N, 3 * Dm), dtype=torch.bfloat16, device=DEV)
    onehot = torch.nn.functional.one_hot(torch.arange(N, device=DEV) % hd, hd).to(torch.bfloat16)
    qkv[:, 2 * Dm:] = onehot.repeat(Bsz, H)
    npad = (N + 15) // 16 * 16
    keep = torch.from_numpy(R.keep_mask(seed, site, p, 0, Bsz * H * N * npad).reshape(Bsz, H, N, npad)[..., :N]).to(DEV)
    want = torch.einsum("bhqk,kd->bqhd", keep.float(), onehot.float()).reshape(Bsz * N, Dm) * (R.inv_keep(p) / N)
    for eng in (L.ENGINE_SIMT, L.ENGINE_TCGEN05):
        o = torch.empty((Bsz * N, Dm), dtype=torch.bfloat16, device=DEV)
        ops.attn_fwd(eng, L.BF16, qkv, o, torch.empty((Bsz, H, N), device=DEV), Bsz, N, H, hd, (seed, site, p))
        # counts are small integers times inv_keep / N: a single flipped mask bit changes an entry by >= 1 / N
        assert (o.float() - want).abs().max().item() < 0.25 * R.inv_keep(p) / N, eng


def test_embed_backward_pieces():
    Bsz, Kp, Fp, Tp, D = 3, 2, 3, 4, 64
    n = Kp * Fp * Tp
    g0 = _rand((Bsz, n + 1, D), L.F32, 80)
    gtok = torch.empty((Bsz * n, D), device=DEV)
    R, dcls = torch.empty((n, D), device=DEV), torch.empty(D, device=DEV)
    ops.embed_bwd_prep(g0, Bsz, n, D, None, gtok, L.F32, R, dcls)
    assert rel_err(gtok, g0[:, 1:].reshape(Bsz * n, D)) < 1e-7
    assert rel_err(R, g0[:, 1:].double().sum(0)) < 1e-6
    assert rel_err(dcls, g0[:, 0].double().sum(0)) < 1e-6
    dpk, dpf, dpt, db = (torch.empty((Kp, D), device=DEV), torch.empty((Fp, D), device=DEV),
                         torch.empty((Tp, D), device=DEV), torch.empty(D, device=DEV))
    ops.pos_grad_reduce(R, Kp, Fp, Tp, D, dpk, dpf, dpt, db)
    R4 = R.double().reshape(Kp, Fp, Tp, D)
    assert rel_err(dpk, R4.sum((1, 2))) < 1e-6 and rel_err(dpf, R4.sum((0, 2))) < 1e-6
    assert rel_err(dpt, R4.sum((0, 1))) < 1e-6 and rel_err(db, R4.sum((0, 1, 2))) < 1e-6


def test_adamw_matches_torch():
    n = 10007
    p0, g = _rand((n,), L.F32, 90), _rand((n,), L.F32, 91)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-4, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        ref.grad = g.clone()
        opt.step()
        ops.adamw(p, g, m, v, 3e-4, 0.9, 0.999, 1e-8, 0.01, step)
    assert rel_err(p, ref.detach()) < 1e-6


def test_attention_probs_rows_sum_to_one():
    Bsz, N, H, hd = 2, 65, 2, 64
    qkv = _rand((Bsz * N, 3 * H * hd), L.F32, 95)
    probs = torch.empty((Bsz, H, N, N), device=DEV)
    ops.attn_probs(L.F32, qkv, probs, Bsz, N, H, hd)
    t = qkv.double().reshape(Bsz, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    ref = torch.softmax((t[0] @ t[1].transpose(-2, -1)) * hd ** -0.5, -1)
    assert rel_err(probs, ref) < 1e-5
    assert torch.allclose(probs.sum(-1), torch.ones_like(probs.sum(-1)), atol=1e-5)

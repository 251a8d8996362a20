"""Shapes of the BASELINE.json configs at reduced depth/batch (B200 only).

* Directly against the fp64 oracle (the restatement pinned to the real reference): the full configs[1] token count
  (8x128x256 -> N = 2049, D384/H6), the large variant's width at that token count (D768/H12, N = 2049) and the
  real-data shape 8x64x488 -> N = 1953.  The oracle's (B,H,N,N) fp64 score tensors are 0.2-0.4 GB
  each at batch 1 and fit a B200 easily.
* The long-sequence variant (32x128x512 -> N = 16385) would need 13 GB per fp64 score tensor and several of them
  alive for backward; there the bf16 tensor-core path is compared with the on-device fp32 verification path, which the
  golden-vector, fp64-oracle and train-mode tests pin to the reference.
Gates (north_star): logits and gradients within 2e-2 relative; every parameter tensor with more than 4096 elements is
held to 2e-2 individually."""
import pytest
import torch

import neural_vit_b200 as nv
from oracle import vit_oracle as O
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    ("large_d768_h12_n513", dict(n_trials=8, freq_size=64, time_size=128, embed_dim=768, n_heads=12, n_layers=2), 2),
    ("small_n2049", dict(n_trials=8, freq_size=128, time_size=256, embed_dim=384, n_heads=6, n_layers=1), 2),
    ("longseq_n16385", dict(n_trials=32, freq_size=128, time_size=512, embed_dim=384, n_heads=6, n_layers=1), 1),
    ("base_d512_h8_n257", dict(n_trials=4, freq_size=64, time_size=128, embed_dim=512, n_heads=8, n_layers=2), 3),
]


@pytest.mark.parametrize("tag,kw,batch", CASES, ids=[c[0] for c in CASES])
def test_bf16_path_matches_fp32_path(tag, kw, batch):
    kw = dict(kw, dropout=0.0, attention_dropout=0.0, drop_path=0.0)
    cfg = nv.Temporal3DViTConfig(**kw)
    params = O.random_params(O.config_from(cfg), seed=21)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(batch, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g).to(DEV)
    y = torch.randint(0, 2, (batch,), generator=g).to(DEV)
    out = {}
    for precision in ("fp32", "bf16"):
        m = nv.Temporal3DViT(cfg, precision=precision)
        m.load_state_dict(params)
        m.to(DEV).train()
        logits = m(x)
        loss = torch.nn.functional.cross_entropy(logits, y)
        loss.backward()
        out[precision] = (logits.detach(), {k: p.grad.detach().clone() for k, p in m.named_parameters()})
        del m
    l32, g32 = out["fp32"]
    l16, g16 = out["bf16"]
    assert torch.isfinite(l16).all()
    assert rel_err(l16, l32) < 2e-2
    flat16 = torch.cat([g16[k].double().flatten() for k in g32])
    flat32 = torch.cat([g32[k].double().flatten() for k in g32])
    assert rel_err(flat16, flat32) < 2e-2
    worst = max((rel_err(g16[k], g32[k]), k) for k in g32 if g32[k].numel() > 4096)
    assert worst[0] < 2e-2, worst


ORACLE_CASES = [
    ("c2_small_n2049_d384_h6", dict(n_trials=8, freq_size=128, time_size=256, embed_dim=384, n_heads=6, n_layers=2), 1),
    ("c5_large_n2049_d768_h12", dict(n_trials=8, freq_size=128, time_size=256, embed_dim=768, n_heads=12, n_layers=1), 1),
    ("realdata_8x64x488_n1953", dict(n_trials=8, freq_size=64, time_size=488, embed_dim=384, n_heads=6, n_layers=1), 2),
]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,kw,batch", ORACLE_CASES, ids=[c[0] for c in ORACLE_CASES])
def test_baseline_shapes_against_fp64_oracle(tag, kw, batch, precision):
    kw = dict(kw, dropout=0.0, attention_dropout=0.0, drop_path=0.0)
    cfg = nv.Temporal3DViTConfig(**kw)
    params = O.random_params(O.config_from(cfg), seed=31)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(batch, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g).to(DEV)
    y = torch.randint(0, 2, (batch,), generator=g).to(DEV)
    m = nv.Temporal3DViT(cfg, precision=precision)
    m.load_state_dict(params)
    m.to(DEV).train()
    logits = m(x)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    del m
    p64 = {k: v.double().to(DEV) for k, v in params.items()}
    rl, rloss, rg = O.loss_and_grads(x.double(), y, p64, O.config_from(cfg))
    if precision == "fp32":
        assert abs(float(loss) - float(rloss)) <= 1e-4
        assert rel_err(logits, rl) < 1e-4
        worst = max((rel_err(grads[k], rg[k]), k) for k in rg)
        assert worst[0] < 1e-3, worst
        return
    assert rel_err(logits, rl) < 2e-2
    flat = torch.cat([grads[k].double().flatten() for k in rg])
    flat_ref = torch.cat([rg[k].double().flatten() for k in rg])
    assert rel_err(flat, flat_ref) < 2e-2
    errs = sorted(((rel_err(grads[k], rg[k]), k, grads[k].numel()) for k in rg), reverse=True)
    for e, k, n in errs:
        assert e < (2e-2 if n > 4096 else 4e-2), errs[:6]


def test_batch_of_one_and_odd_batch():
    cfg = nv.Temporal3DViTConfig(n_trials=4, freq_size=32, time_size=64, embed_dim=128, n_heads=2, n_layers=1,
                                 dropout=0.0, attention_dropout=0.0, drop_path=0.0)
    params = O.random_params(O.config_from(cfg), seed=3)
    m = nv.Temporal3DViT(cfg)
    m.load_state_dict(params)
    m.to(DEV).eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(5, 4, 32, 64, generator=g).to(DEV)
    with torch.no_grad():
        full = m(x)
        one = torch.cat([m(x[i:i + 1]) for i in range(5)])
    assert rel_err(one, full) < 1e-5       # samples are independent (batch-sharded data parallelism relies on it)
    with pytest.raises(ValueError):
        m(torch.zeros(2, 4, 32, 60, device=DEV))


@pytest.mark.parametrize("B,N,H,drop", [(32, 2049, 6, (1, 2, 0.1)), (8, 1000, 12, (3, 4, 0.3)), (32, 2049, 6, None)])
def test_attention_kernels_are_race_free(B, N, H, drop):
    """dK / dV and the forward output are produced without atomics, so repeated launches on the same inputs must be
    bit-identical at full occupancy (every CTA slot busy, all pipeline stages recycled many times); a missing
    barrier in the TMEM / shared-memory hand-offs shows up here as a flipped bit.  dQ goes through fp32 atomics and
    is compared numerically."""
    from neural_vit_b200 import _lib as L, ops
    hd = 64
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn(B * N, 3 * D, device="cuda", generator=g).bfloat16()
    dout = torch.randn(B * N, D, device="cuda", generator=g).bfloat16()
    outs, grads = [], []
    for _ in range(4):
        out = torch.empty(B * N, D, dtype=torch.bfloat16, device="cuda")
        lse = torch.empty(B, H, N, device="cuda")
        ops.attn_fwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, lse, B, N, H, hd, drop)
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(L.ENGINE_TCGEN05, L.BF16, qkv, out, dout, lse, dqkv, B, N, H, hd, drop)
        outs.append(out)
        grads.append(dqkv)
    for o, gr in zip(outs[1:], grads[1:]):
        assert torch.equal(o, outs[0])
        assert torch.equal(gr[:, D:], grads[0][:, D:])
        assert rel_err(gr[:, :D], grads[0][:, :D]) < 1e-3

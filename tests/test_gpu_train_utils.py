"""B200 tests of the pieces right around the hot path (SURVEY.md section 8 rows e, f-1 ... f-4): fused AdamW with
operand-shadow re-cast, gradient sink (backward kernels accumulate into flat buckets), shadow staleness, device-side
loss + metrics, the pinned double-buffered input feed and the tensor-core attention-map path."""
import numpy as np
import pytest
import torch

import neural_vit_b200 as nv
from neural_vit_b200 import _lib as L, ops
from oracle import vit_oracle as O
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

KW = dict(n_trials=4, freq_size=32, time_size=64, embed_dim=128, n_heads=2, n_layers=2, dropout=0.0,
          attention_dropout=0.0, drop_path=0.0)


def _data(cfg, batch, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g).to(DEV)
    y = torch.randint(0, 2, (batch,), generator=g).to(DEV)
    return x, y


def _model(kw, precision, seed=7, randomise=True):
    cfg = nv.Temporal3DViTConfig(**kw)
    torch.manual_seed(seed)
    m = nv.Temporal3DViT(cfg, precision=precision)
    if randomise:
        m.load_state_dict(O.random_params(O.config_from(cfg), seed=seed))
    return m.to(DEV).train()


@pytest.mark.parametrize("layer_scale", [1e-4, 0.0])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_grad_sink_equals_autograd_gradients(precision, layer_scale):
    """Gradients written by the backward kernels straight into the flat buckets == the gradients autograd collects;
    two passes without a consumer accumulate; a consumed sink is cleared by the next forward."""
    kw = dict(KW, layer_scale_init=layer_scale)
    a, b = _model(kw, precision), _model(kw, precision)
    x, y = _data(a.config, 4)
    torch.nn.functional.cross_entropy(a(x), y).backward()
    sink = b.attach_grad_sink(bucket_mb=0.25)
    assert len(sink.buckets) >= 2
    torch.nn.functional.cross_entropy(b(x), y).backward()
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert pb.grad is not None and pb.grad.data_ptr() == sink.view(pb).data_ptr(), k
        assert rel_err(pb.grad, pa.grad) < (1e-5 if precision == "fp32" else 2e-3), k
    g1 = {k: p.grad.clone() for k, p in b.named_parameters()}
    torch.nn.functional.cross_entropy(b(x), y).backward()          # no consumer in between -> accumulates
    for k, p in b.named_parameters():
        assert rel_err(p.grad, 2 * g1[k]) < 1e-4, k
    sink.consumed = True                                            # what step() / finish() do
    torch.nn.functional.cross_entropy(b(x), y).backward()
    for k, p in b.named_parameters():
        assert rel_err(p.grad, g1[k]) < 1e-4, k
    b.zero_grad(set_to_none=True)                                   # a stock zero_grad is honoured too
    torch.nn.functional.cross_entropy(b(x), y).backward()
    for k, p in b.named_parameters():
        assert p.grad is not None and rel_err(p.grad, g1[k]) < 1e-4, k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_adamw_matches_torch_adamw_and_refreshes_shadows(precision):
    """10 steps of FusedAdamW == 10 steps of torch.optim.AdamW (train.py:154-156,227) on parameters; after every step
    the adopted bf16 shadows equal a fresh cast of the fp32 masters, so no forward ever runs on stale operands."""
    a, b = _model(KW, precision, randomise=False), _model(KW, precision, randomise=False)
    opt_a = torch.optim.AdamW(a.parameters(), lr=3e-3, weight_decay=0.05)
    opt_b = nv.FusedAdamW(b.parameters(), lr=3e-3, weight_decay=0.05, model=b)
    assert len(opt_b.param_groups) == 1 and opt_b.param_groups[0]["lr"] == 3e-3
    crit = torch.nn.CrossEntropyLoss(weight=torch.tensor([0.8, 1.3], device=DEV), label_smoothing=0.05)
    for step in range(10):
        x, y = _data(a.config, 8, seed=step)
        for m, opt in ((a, opt_a), (b, opt_b)):
            opt.zero_grad()
            crit(m(x), y).backward()
            opt.step()
        if precision == "bf16":
            sh = b._shadows
            for i, blk in enumerate(b.blocks):
                for name, w, gamma in (("qkv", blk.attn.qkv.weight, None), ("proj", blk.attn.proj.weight, blk.ls1.gamma),
                                       ("fc1", blk.mlp.fc1.weight, None), ("fc2", blk.mlp.fc2.weight, blk.ls2.gamma)):
                    w_sh, wt_sh = sh.peek((i, name))
                    assert torch.equal(w_sh, w.detach().bfloat16()), (step, i, name)
                    want_t = (w.detach() if gamma is None else gamma.detach()[:, None] * w.detach()).t().bfloat16()
                    assert torch.equal(wt_sh, want_t), (step, i, name)
    tol = 2e-4 if precision == "fp32" else 3e-2      # bf16: the two runs see differently-ordered bf16 rounding
    D = a.config.embed_dim
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        if k.endswith("attn.qkv.bias"):
            # softmax is invariant to a shift of the scores along the key axis, so the gradient of the K third of
            # the qkv bias is exactly zero in exact arithmetic; what each run sees there is rounding noise that Adam
            # normalises to O(lr) steps of random sign.  Compare the Q and V thirds.
            pa, pb = torch.cat([pa[:D], pa[2 * D:]]), torch.cat([pb[:D], pb[2 * D:]])
        assert rel_err(pb, pa) < tol, k
    # same loss on a fresh batch => the forward really uses the updated (adopted) operands
    x, y = _data(a.config, 8, seed=99)
    with torch.no_grad():
        assert rel_err(b(x), a(x)) < (1e-3 if precision == "fp32" else 5e-2)
    sd = opt_b.state_dict()
    assert len(sd["state"]) == len(list(b.parameters())) and sd["param_groups"][0]["weight_decay"] == 0.05


def test_fused_adamw_plain_parameter_list_matches_torch():
    torch.manual_seed(0)
    net_a = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.GELU(), torch.nn.Linear(64, 3)).to(DEV)
    net_b = torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.GELU(), torch.nn.Linear(64, 3)).to(DEV)
    net_b.load_state_dict(net_a.state_dict())
    oa = torch.optim.AdamW(net_a.parameters(), lr=1e-2, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1)
    ob = nv.FusedAdamW(net_b.parameters(), lr=1e-2, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1)
    x = torch.randn(16, 37, device=DEV)
    for _ in range(25):
        for net, opt in ((net_a, oa), (net_b, ob)):
            opt.zero_grad()
            net(x).square().mean().backward()
            opt.step()
    for pa, pb in zip(net_a.parameters(), net_b.parameters()):
        assert rel_err(pb, pa) < 1e-5


def test_raw_pointer_updates_never_leave_stale_shadows():
    """ops.adamw writes parameters through raw pointers; the next forward must see them."""
    m = _model(KW, "bf16")
    x, _ = _data(m.config, 2)
    with torch.no_grad():
        before = m(x).clone()
        w = m.blocks[0].mlp.fc1.weight
        g = torch.ones_like(w)
        ops.adamw(w.view(-1), g.view(-1), torch.zeros_like(w).view(-1), torch.zeros_like(w).view(-1), 0.5, 0.9, 0.999,
                  1e-8, 0.0, 1)
        after = m(x).clone()
        assert rel_err(after, before) > 1e-3
        w_sh, _ = m._shadows.peek((0, "fc1"))
        assert torch.equal(w_sh, w.detach().bfloat16())
        w.data.add_(0.25)                 # .data has its own version counter: only the explicit hook helps here
        m.invalidate_shadows()
        m(x)
        assert torch.equal(m._shadows.peek((0, "fc1"))[0], w.detach().bfloat16())


@pytest.mark.parametrize("B,C,ls,weighted", [(256, 2, 0.05, True), (7, 2, 0.0, False), (33, 5, 0.1, True)])
def test_device_cross_entropy_matches_torch(B, C, ls, weighted):
    g = torch.Generator().manual_seed(B)
    logits = (3 * torch.randn(B, C, generator=g)).to(DEV).requires_grad_(True)
    labels = torch.randint(0, C, (B,), generator=g).to(DEV)
    w = (0.5 + torch.rand(C, generator=g)).to(DEV) if weighted else None
    metrics = nv.DeviceMetrics(DEV, capacity=8)
    crit = nv.CrossEntropyLoss(weight=w, label_smoothing=ls, metrics=metrics)
    loss = crit(logits, labels)
    (2.0 * loss).backward()
    ref_logits = logits.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(ref_logits, labels, weight=w, label_smoothing=ls)
    (2.0 * ref).backward()
    assert abs(float(loss) - float(ref)) < 1e-5
    assert rel_err(logits.grad, ref_logits.grad) < 1e-5
    # second batch through the no-grad (evaluate) path, then ONE host sync for the epoch metrics
    logits2 = (3 * torch.randn(B, C, generator=g)).to(DEV)
    labels2 = torch.randint(0, C, (B,), generator=g).to(DEV)
    with torch.no_grad():
        loss2 = crit(logits2, labels2)
    ref2 = torch.nn.functional.cross_entropy(logits2, labels2, weight=w, label_smoothing=ls)
    assert abs(float(loss2) - float(ref2)) < 1e-5
    out = metrics.compute()
    all_logits, all_labels = torch.cat([logits.detach(), logits2]), torch.cat([labels, labels2])
    assert out["count"] == 2 * B
    assert abs(out["loss"] - (float(ref) + float(ref2)) / 2) < 1e-5                      # train.py:229,237
    assert abs(out["acc"] - float((all_logits.argmax(1) == all_labels).float().mean())) < 1e-6
    if C == 2:
        from sklearn.metrics import roc_auc_score
        probs = torch.softmax(all_logits, 1)[:, 1].cpu().numpy()
        assert abs(out["auc"] - roc_auc_score(all_labels.cpu().numpy(), probs)) < 1e-6


def test_device_prefetcher_delivers_every_batch_in_order():
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(6, 4, 32, 64, generator=g), torch.randint(0, 2, (6,), generator=g)) for _ in range(7)]
    batches[3] = (batches[3][0].pin_memory(), batches[3][1].pin_memory())     # already-pinned batches are used in place
    pf = nv.DevicePrefetcher(batches, DEV)
    seen = 0
    for (xd, yd), (xh, yh) in zip(pf, batches):
        assert xd.is_cuda and torch.equal(xd.cpu(), xh) and torch.equal(yd.cpu(), yh)
        (xd * 2).sum().item()          # consumer work on the compute stream
        seen += 1
    assert seen == 7 and pf.h2d_bytes == 7 * (6 * 4 * 32 * 64 * 4 + 6 * 8)
    assert len(list(nv.DevicePrefetcher([], DEV))) == 0


@pytest.mark.parametrize("kw,batch", [
    (dict(n_trials=4, freq_size=32, time_size=64, embed_dim=128, n_heads=2, n_layers=2), 2),       # N = 65
    (dict(n_trials=4, freq_size=64, time_size=128, embed_dim=192, n_heads=3, n_layers=1), 1),      # N = 257
])
def test_attention_maps_tensor_core_path(kw, batch):
    """get_attention_maps on the bf16 path: P = exp(scale QK^T - lse) materialised by the tcgen05 GEMM epilogue."""
    kw = dict(kw, dropout=0.0, attention_dropout=0.0, drop_path=0.0)
    cfg = nv.Temporal3DViTConfig(**kw)
    params = O.random_params(O.config_from(cfg), seed=4)
    x, _ = _data(cfg, batch, seed=2)
    ref = O.attention_maps(x.double(), {k: v.double().to(DEV) for k, v in params.items()}, O.config_from(cfg))
    for precision, tol in (("fp32", 1e-4), ("bf16", 3e-2)):
        m = nv.Temporal3DViT(cfg, precision=precision)
        m.load_state_dict(params)
        m.to(DEV).eval()
        maps = m.get_attention_maps(x)
        assert len(maps) == cfg.n_layers
        for got, want in zip(maps, ref):
            assert got.shape == want.shape and got.dtype == torch.float32
            assert rel_err(got, want) < tol
            assert float((got.sum(-1) - 1).abs().max()) < (1e-4 if precision == "fp32" else 2e-2)

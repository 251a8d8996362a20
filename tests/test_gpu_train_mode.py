"""Train-mode (p > 0) parity of the WHOLE model against the oracle with the library's own masks.

ATen's dropout stream cannot be reproduced, but the library's generator is a pure function of
(seed, site, element) and oracle/dropout_ref.py restates it bit for bit.  So: run one train-mode step of the drop-in,
rebuild every site's keep mask and the DropPath draws with tests/masks.py, feed them to the fp64 oracle
(``vit_oracle.loss_and_grads(masks=...)``) and compare logits, loss and every parameter gradient.  This covers the
sites the kernel-level mask tests do not: pos_drop in the patch-embed epilogue / ``cls_rows`` / ``embed_bwd_prep``,
head dropout, DropPath row scaling, and the site-id and element-index conventions shared by forward and backward.
Gates (BASELINE.json north_star): fp32 path |loss - ref| <= 1e-4; bf16 path logits and gradients within 2e-2.
"""
import pytest
import torch

import neural_vit_b200 as nv
from oracle import vit_oracle as O
from tests import masks as MK
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # tag, config, batch, torch seed
    ("n65_d128_L3_p0.1", dict(n_trials=4, freq_size=32, time_size=64, embed_dim=128, n_heads=2, n_layers=3,
                              dropout=0.1, attention_dropout=0.1, drop_path=0.1), 3, 11),
    ("n129_d192_L2_mixed", dict(n_trials=4, freq_size=32, time_size=128, embed_dim=192, n_heads=3, n_layers=2,
                                dropout=0.2, attention_dropout=0.1, drop_path=0.5), 4, 5),
    ("n65_d128_nols", dict(n_trials=4, freq_size=32, time_size=64, embed_dim=128, n_heads=2, n_layers=2,
                           dropout=0.1, attention_dropout=0.1, drop_path=0.3, layer_scale_init=0.0), 4, 2),
]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,kw,batch,seed", CASES, ids=[c[0] for c in CASES])
def test_train_mode_matches_oracle_with_library_masks(tag, kw, batch, seed, precision):
    cfg = nv.Temporal3DViTConfig(**kw)
    params = O.random_params(O.config_from(cfg), seed=13)
    m = nv.Temporal3DViT(cfg, precision=precision)
    m.load_state_dict(params)
    m.to(DEV).train()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, cfg.n_trials, cfg.freq_size, cfg.time_size, generator=g).to(DEV)
    y = torch.randint(0, 2, (batch,), generator=g).to(DEV)
    cw = torch.tensor([0.8, 1.3], device=DEV)
    torch.manual_seed(seed)
    logits = m(x)
    loss = torch.nn.functional.cross_entropy(logits, y, weight=cw, label_smoothing=0.05)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}

    mk = MK.build_masks(m, batch, device=DEV)
    if kw["drop_path"] >= 0.3:     # the DropPath branch must actually be exercised by these draws
        assert any(float(v.min()) == 0.0 for k, v in mk.items() if "drop_path" in k)
    for k, v in mk.items():        # masks realise the requested rates
        if "drop_path" not in k and v.numel() > 20000:
            want = 1.0 - MK.effective_rate(kw["attention_dropout"] if "attn_drop" in k else kw["dropout"])
            assert abs(float(v.mean()) - want) < 0.01, k
    ocfg = MK.oracle_config(cfg)
    p64 = {k: v.double().to(DEV) for k, v in params.items()}
    rl, rloss, rg = O.loss_and_grads(x.double(), y, p64, ocfg, class_weight=cw.double(), label_smoothing=0.05,
                                     masks={k: v.double() for k, v in mk.items()})
    if precision == "fp32":
        assert abs(float(loss) - float(rloss)) <= 1e-4
        assert rel_err(logits, rl) < 1e-4
        worst = max((rel_err(grads[k], rg[k]), k) for k in rg)
        assert worst[0] < 5e-4, worst
    else:
        assert rel_err(logits, rl) < 2e-2
        flat = torch.cat([grads[k].double().flatten() for k in rg])
        flat_ref = torch.cat([rg[k].double().flatten() for k in rg])
        assert rel_err(flat, flat_ref) < 2e-2
        errs = sorted(((rel_err(grads[k], rg[k]), k, grads[k].numel()) for k in rg), reverse=True)
        # D = 128 with O(1) LayerScale sits at the bf16-activation error floor: 4e-2 per tensor
        # there and for vectors of <= 4096 elements; the realistic width (D = 192) is held to 2e-2 per matrix
        for e, k, n in errs:
            assert e < (2e-2 if (cfg.embed_dim >= 192 and n > 4096) else 4e-2), errs[:6]


def test_same_seed_replays_identical_draws():
    cfg = nv.Temporal3DViTConfig(**CASES[0][1])
    m = nv.Temporal3DViT(cfg, precision="bf16").to(DEV).train()
    x = torch.randn(2, cfg.n_trials, cfg.freq_size, cfg.time_size, device=DEV)
    torch.manual_seed(3)
    a = m(x)
    d1 = m.last_draws["seed"]
    torch.manual_seed(3)
    b = m(x)
    assert m.last_draws["seed"] == d1 and torch.equal(a, b)
    c = m(x)
    assert m.last_draws["seed"] != d1 and not torch.equal(a, c)

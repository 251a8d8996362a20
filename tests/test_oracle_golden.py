"""Pin the oracle (oracle/vit_oracle.py) against fixtures generated from the REAL reference."""
import numpy as np
import pytest
import torch

from oracle import vit_oracle as O
from tests.conftest import load_golden, rel_err

CASES = ["g1_eval_d64", "g2_train_d128", "g3_train_nodrop_nols"]


def _run(name):
    g = load_golden(name)
    cfg = O.config_from(g["cfg"])
    x = torch.from_numpy(g["x"])
    y = torch.from_numpy(g["y"])
    masks = g["mask"] if int(g["train_mode"]) else None
    cw = torch.from_numpy(g["class_weight"]) if "class_weight" in g else None
    ls = float(g["label_smoothing"])
    logits, loss, grads = O.loss_and_grads(x, y, g["param"], cfg, class_weight=cw,
                                           label_smoothing=ls, masks=masks)
    return g, cfg, logits, loss, grads


@pytest.mark.parametrize("name", CASES)
def test_logits_and_loss_match_reference(name):
    g, cfg, logits, loss, grads = _run(name)
    assert rel_err(logits, torch.from_numpy(g["logits"])) < 2e-6
    assert abs(float(loss) - float(g["loss"])) < 2e-6


@pytest.mark.parametrize("name", CASES)
def test_param_grads_match_reference(name):
    g, cfg, logits, loss, grads = _run(name)
    assert set(grads) == set(g["grad"])
    for k, ref in g["grad"].items():
        assert grads[k].shape == ref.shape, k
        assert rel_err(grads[k], ref) < 2e-5, k


@pytest.mark.parametrize("name", CASES)
def test_intermediates_match_reference(name):
    g = load_golden(name)
    cfg = O.config_from(g["cfg"])
    masks = g["mask"] if int(g["train_mode"]) else None
    taps = {}
    O.forward(torch.from_numpy(g["x"]), g["param"], cfg, masks=masks, taps=taps)
    checked = 0
    for k, ref in g["tap"].items():
        assert rel_err(taps[k], ref) < 2e-6, k
        checked += 1
    assert checked >= 6 * cfg.n_layers


def test_attention_maps_match_reference():
    g = load_golden("g1_eval_d64")
    cfg = O.config_from(g["cfg"])
    maps = O.attention_maps(torch.from_numpy(g["x"]), g["param"], cfg)
    assert len(maps) == cfg.n_layers
    assert rel_err(maps[0], torch.from_numpy(g["attn_map.0"])) < 2e-6
    assert rel_err(maps[-1], torch.from_numpy(g["attn_map.last"])) < 2e-6
    # rows of a softmax sum to one (size-independent property)
    assert torch.allclose(maps[0].sum(-1), torch.ones_like(maps[0].sum(-1)), atol=1e-5)


def test_im2col_is_conv3d():
    """Conv3d with kernel == stride equals im2col + GEMM (SURVEY.md 8c probe)."""
    cfg = O.OracleConfig(n_trials=4, freq_size=16, time_size=24, embed_dim=32, n_heads=1)
    torch.manual_seed(0)
    x = torch.randn(2, 4, 16, 24)
    w = torch.randn(32, 1, 2, 8, 8)
    b = torch.randn(32)
    ref = torch.nn.functional.conv3d(x[:, None], w, b, stride=(2, 8, 8)).flatten(2).transpose(1, 2)
    got = O.tubelet_im2col(x, cfg) @ w.reshape(32, -1).T + b
    assert rel_err(got, ref) < 1e-6


def test_param_shapes_cover_state_dict():
    g = load_golden("g1_eval_d64")
    cfg = O.config_from(g["cfg"])
    shapes = O.param_shapes(cfg)
    assert list(shapes) and set(shapes) == set(g["param"])
    for k, s in shapes.items():
        assert tuple(g["param"][k].shape) == s


def test_drop_path_rates_block0_is_zero():
    cfg = O.OracleConfig()
    r = O.drop_path_rates(cfg)
    assert r[0] == 0.0 and abs(r[-1] - cfg.drop_path) < 1e-7 and len(r) == cfg.n_layers

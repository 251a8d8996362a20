"""CPU-side checks: drop-in boundary (config, state_dict layout, initialisation), loud failure
without CUDA, and that the C-ABI library exports every symbol include/tvit.h declares."""
import ctypes
import os
import re
from dataclasses import asdict

import numpy as np
import pytest
import torch

import neural_vit_b200 as nv
from neural_vit_b200 import _lib
from oracle import vit_oracle as O
from tests.conftest import GOLDEN_DIR, ROOT, load_golden


def test_config_matches_reference_fields():
    g = load_golden("g1_eval_d64")
    ours = asdict(nv.Temporal3DViTConfig())
    assert list(ours) == list(g["cfg"])          # same field names, same order
    c = nv.Temporal3DViTConfig()
    assert (c.n_trials, c.freq_size, c.time_size, c.embed_dim, c.n_heads, c.n_layers) == (8, 64, 128, 384, 6, 8)
    assert c.n_patches == 4 * 8 * 16 and c.patch_dim == 128
    assert set(nv.CONFIGS) == {"tiny", "small", "base"}
    assert (nv.CONFIGS["base"].embed_dim, nv.CONFIGS["base"].n_heads, nv.CONFIGS["base"].n_layers) == (512, 8, 12)


@pytest.mark.parametrize("name", ["g1_eval_d64", "g2_train_d128", "g3_train_nodrop_nols"])
def test_state_dict_layout_matches_reference(name):
    g = load_golden(name)
    m = nv.Temporal3DViT(nv.Temporal3DViTConfig(**g["cfg"]))
    sd = m.state_dict()
    assert set(sd) == set(g["param"])
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(g["param"][k].shape), k
        assert v.dtype == torch.float32
    # a reference checkpoint loads strictly, and saves back identically
    missing = m.load_state_dict(g["param"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(sd) == list(O.param_shapes(O.config_from(g["cfg"])))   # registration order


@pytest.mark.parametrize("tag,kw", [
    ("tiny", dict(embed_dim=192, n_heads=3, n_layers=4)),
    ("small_8x128x256", dict(embed_dim=384, n_heads=6, n_layers=8, n_trials=8, freq_size=128, time_size=256)),
])
def test_same_seed_same_initial_weights_as_reference(tag, kw):
    z = np.load(os.path.join(GOLDEN_DIR, "init_checksums.npz"))
    torch.manual_seed(1234)
    m = nv.Temporal3DViT(nv.Temporal3DViTConfig(**kw))
    sd = m.state_dict()
    names = [str(s) for s in z[tag + ".names"]]
    assert list(sd) == names
    sums = z[tag + ".sums"]
    for i, k in enumerate(names):
        v = sd[k].double()
        assert v.numel() == int(sums[i, 2])
        assert abs(float(v.sum()) - sums[i, 0]) <= 1e-9 * max(1.0, abs(sums[i, 0])), k
        assert abs(float((v ** 2).sum()) - sums[i, 1]) <= 1e-9 * max(1.0, abs(sums[i, 1])), k


def test_constructor_errors_match_reference():
    with pytest.raises(ValueError, match="n_trials"):
        nv.Temporal3DViT(nv.Temporal3DViTConfig(n_trials=7))
    with pytest.raises(ValueError, match="freq_size"):
        nv.Temporal3DViT(nv.Temporal3DViTConfig(freq_size=60))
    with pytest.raises(ValueError, match="time_size"):
        nv.Temporal3DViT(nv.Temporal3DViTConfig(time_size=100))
    with pytest.raises(ValueError, match="precision"):
        nv.Temporal3DViT(nv.Temporal3DViTConfig(), precision="fp8")


def test_no_layer_scale_and_block0_drop_path():
    m = nv.Temporal3DViT(nv.Temporal3DViTConfig(layer_scale_init=0.0, n_layers=3, drop_path=0.2))
    assert not any("gamma" in k for k in m.state_dict())
    rates = [b.drop_path_rate for b in m.blocks]
    assert rates[0] == 0.0 and abs(rates[-1] - 0.2) < 1e-6


def test_forward_on_cpu_fails_loudly():
    m = nv.Temporal3DViT(nv.CONFIGS["tiny"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 8, 64, 128))


def test_checkpoint_roundtrip_reference_layout(tmp_path):
    m = nv.Temporal3DViT(nv.CONFIGS["tiny"])
    ckpt = {"model_state": m.state_dict(), "config": asdict(m.config)}   # train.py:268-271
    path = tmp_path / "final.pt"
    torch.save(ckpt, path)
    back = torch.load(path)
    m2 = nv.Temporal3DViT(nv.Temporal3DViTConfig(**back["config"]))
    m2.load_state_dict(back["model_state"])
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k])


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tvit.h")).read()
    declared = set(re.findall(r"\b(tvit_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} missing from libtvit_b200.so"
    assert _lib.load().tvit_version() >= 100


def test_shim_import_path():
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "shim"))
    try:
        sys.modules.pop("temporal_vit", None)
        sys.modules.pop("temporal_vit.models", None)
        sys.modules.pop("temporal_vit.models.model", None)
        mod = importlib.import_module("temporal_vit.models.model")
        assert mod.Temporal3DViT is nv.Temporal3DViT and mod.CONFIGS is nv.CONFIGS
    finally:
        sys.path.remove(os.path.join(ROOT, "shim"))
        for k in [k for k in sys.modules if k.startswith("temporal_vit")]:
            sys.modules.pop(k)

"""CPU-side checks of the pieces around the hot path (SURVEY.md section 8 rows b, e, f): the import shim resolves the
reference's own training package, the rank-sharding sampler partitions an epoch, the AUC routine equals sklearn's,
checkpoints round-trip in the reference layout, and the hook-mode all-reduce supports gradient accumulation."""
import os
import subprocess
import sys
from dataclasses import asdict

import numpy as np
import pytest
import torch

import neural_vit_b200 as nv
from tests import refrun
from tests.conftest import ROOT


@pytest.mark.skipif(refrun.reference_root() is None, reason="reference sources not installed (baseline/_ref)")
def test_shim_resolves_reference_training_package():
    """with shim/ ahead of the reference on PYTHONPATH the reference's train module imports and
    binds the B200 model classes (train.py:13), while every other reference module stays the reference's own."""
    ref = refrun.reference_root()
    code = ("import temporal_vit.training.train as t, temporal_vit.training.train_hptune as h, neural_vit_b200 as nv;"
            "import temporal_vit.data.data_loader as d, temporal_vit.models.model as m;"
            "assert t.Temporal3DViT is nv.Temporal3DViT and t.CONFIGS is nv.CONFIGS;"
            "assert t.Temporal3DViTConfig is nv.Temporal3DViTConfig and h.Temporal3DViT is nv.Temporal3DViT;"
            "print(t.__file__); print(d.__file__); print(m.__file__)")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "shim"), ref, ROOT]))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0, r.stderr[-2000:]
    train_file, data_file, model_file = r.stdout.strip().splitlines()[-3:]
    assert train_file.startswith(ref) and data_file.startswith(ref)
    assert model_file.startswith(os.path.join(ROOT, "shim"))


def test_rank_shard_sampler_partitions_the_epoch():
    n, world = 103, 4
    shards = [nv.RankShardSampler(n, r, world, shuffle=True, seed=3) for r in range(world)]
    for s in shards:
        s.set_epoch(2)
    got = [list(s) for s in shards]
    assert len({len(g) for g in got}) == 1 and len(got[0]) == len(shards[0]) == 26
    flat = sorted(i for g in got for i in g)
    assert set(flat) == set(range(n)) and len(flat) == 104           # one wrapped sample pads the last round
    shards[0].set_epoch(3)
    assert list(shards[0]) != got[0]                                  # a new epoch reshuffles
    drop = [list(nv.RankShardSampler(n, r, world, shuffle=False, drop_last=True)) for r in range(world)]
    assert sorted(i for g in drop for i in g) == list(range(100))
    with pytest.raises(ValueError):
        nv.RankShardSampler(10, 4, 4)


def test_roc_auc_matches_sklearn_including_ties():
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(0)
    for k in range(25):
        y = rng.integers(0, 2, 300)
        s = np.round(rng.random(300), 1 + k % 3)                      # coarse rounding -> many ties
        assert abs(nv.roc_auc(y, s) - roc_auc_score(y, s)) < 1e-12
    assert np.isnan(nv.roc_auc(np.ones(5), np.arange(5.0)))           # train.py:101-102: single class -> nan


def test_checkpoint_loader_reads_reference_layout(tmp_path):
    torch.manual_seed(3)
    m = nv.Temporal3DViT(nv.CONFIGS["tiny"])
    path = str(tmp_path / "final.pt")
    nv.save_checkpoint(m, path)
    raw = torch.load(path)
    assert set(raw) == {"model_state", "config"} and raw["config"] == asdict(m.config)     # train.py:268-271
    m2 = nv.load_checkpoint(path, precision="fp32")
    assert m2.precision == "fp32" and m2.config == m.config
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k]), k
    torch.save({"weights": {}}, path)
    with pytest.raises(ValueError):
        nv.load_checkpoint(path)


@pytest.mark.skipif(refrun.reference_root() is None, reason="reference sources not installed (baseline/_ref)")
def test_checkpoint_written_by_the_reference_model_loads(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_model_for_test", os.path.join(refrun.reference_root(), "temporal_vit", "models", "model.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(5)
    rm = ref.Temporal3DViT(ref.Temporal3DViTConfig(embed_dim=128, n_heads=2, n_layers=2))
    path = str(tmp_path / "ref.pt")
    torch.save({"model_state": rm.state_dict(), "config": asdict(rm.config)}, path)
    m = nv.load_checkpoint(path)
    for k, v in rm.state_dict().items():
        assert torch.equal(v, m.state_dict()[k]), k

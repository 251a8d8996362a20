"""North-star gate "validation AUC/accuracy parity after a fixed synthetic run", and SURVEY.md section 8(b)/(a11): the
reference's OWN ``temporal_vit.training.train.train(cfg)`` (build_model, DataLoaders, AdamW, weighted label-smoothed
CE, epoch loop, ``evaluate``, checkpoint writer -- train.py:53-105,108-305) runs UNCHANGED on the drop-in model,
substituted through ``shim/`` on PYTHONPATH.

The golden record tests/golden/train_run.json was produced by tests/golden/make_train_golden.py from the real
reference model on the CPU (same data, same init seed, same order).  Gates: per-epoch val/auc within
0.02, final val/acc within 0.03, for both the fp32 verification path and the bf16 tensor-core path.
"""
import importlib.util
import json
import os

import pytest
import torch

from tests import refrun
from tests.conftest import GOLDEN_DIR

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(refrun.reference_root() is None,
                                 reason="reference sources not installed (baseline/_ref; tools/install_reference.py)")]


def _golden():
    with open(os.path.join(GOLDEN_DIR, "train_run.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_reference_train_runs_unchanged_and_matches_auc_acc(tmp_path, precision):
    gold = _golden()
    assert gold["run"] == json.loads(json.dumps(refrun.RUN)), "golden record was made with a different run spec"
    out = refrun.run_reference_train(str(tmp_path), "dropin", "cuda", precision=precision)
    assert os.path.join("shim", "temporal_vit", "models", "model.py") in out["model_module"]
    ref_epochs = [m for m in gold["metrics"] if "val/auc" in m]
    got_epochs = [m for m in out["metrics"] if "val/auc" in m]
    assert len(got_epochs) == len(ref_epochs) == refrun.RUN["epochs"]
    report = [(r["step"], round(g["val/auc"] - r["val/auc"], 4), round(g["val/acc"] - r["val/acc"], 4),
               round(g["val/loss"] - r["val/loss"], 4)) for g, r in zip(got_epochs, ref_epochs)]
    print(precision, "epoch, d_auc, d_acc, d_loss:", report)
    for g, r in zip(got_epochs, ref_epochs):
        assert abs(g["val/auc"] - r["val/auc"]) <= 0.02, report
        assert abs(g["train/auc"] - r["train/auc"]) <= 0.03, report
    g, r = got_epochs[-1], ref_epochs[-1]
    assert abs(g["val/acc"] - r["val/acc"]) <= 0.03, report
    assert abs(g["train/acc"] - r["train/acc"]) <= 0.03, report
    gt = [m for m in out["metrics"] if "test/auc" in m][0]
    rt = [m for m in gold["metrics"] if "test/auc" in m][0]
    assert abs(gt["test/auc"] - rt["test/auc"]) <= 0.02 and abs(gt["test/acc"] - rt["test/acc"]) <= 0.06
    if precision == "fp32":   # the verification path follows the reference's trajectory closely all the way
        for g, r in zip(got_epochs, ref_epochs):
            assert abs(g["train/loss"] - r["train/loss"]) <= 0.02, report
    # the checkpoint the reference's writer produced from the drop-in loads into the REFERENCE class (train.py:290-295)
    ckpt = torch.load(out["final"], map_location="cpu", weights_only=True)
    spec = importlib.util.spec_from_file_location(
        "ref_model_for_ckpt", os.path.join(refrun.reference_root(), "temporal_vit", "models", "model.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rm = ref.Temporal3DViT(ref.Temporal3DViTConfig(**ckpt["config"]))
    res = rm.load_state_dict(ckpt["model_state"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys

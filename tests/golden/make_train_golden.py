"""Golden record of the fixed synthetic training run, produced by the REAL reference (build container only).

    python tests/golden/make_train_golden.py

Runs the unmodified reference ``temporal_vit.training.train.train(cfg)`` (train.py:108-305: build_model, AdamW,
class-weighted label-smoothed CE, the epoch loop :207-257 and ``evaluate`` :77-105) with the reference model on the
CPU, on the synthetic parquet set of tests/refrun.py, and stores the per-epoch metrics it logged
(train/val loss, acc, auc + the test metrics) in tests/golden/train_run.json together with the run spec.
tests/test_gpu_train_parity.py trains the drop-in model through the same ``train()`` on a B200 and holds
val/auc to 0.02 and val/acc to 0.03 of this record (BASELINE.json north_star; SURVEY.md section 8d parity gates).
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import refrun  # noqa: E402


def main():
    with tempfile.TemporaryDirectory() as d:
        out = refrun.run_reference_train(d, "reference", "cpu")
    assert "baseline/_ref" in out["model_module"] or "/root/reference" in out["model_module"], out["model_module"]
    import torch
    rec = {"run": refrun.RUN, "torch": torch.__version__, "model_module": "reference temporal_vit/models/model.py",
           "metrics": out["metrics"]}
    path = os.path.join(HERE, "train_run.json")
    with open(path, "w") as fh:
        json.dump(rec, fh, indent=1)
    for m in out["metrics"]:
        print(m)
    print("wrote", path)


if __name__ == "__main__":
    main()

"""Test infrastructure: drive the reference's OWN ``temporal_vit.training.train.train(cfg)`` (train.py:108-305) on a
fixed synthetic parquet data set, either with the reference model (CPU, ``model="reference"``) or with the drop-in
B200 model substituted through ``shim/`` (``model="dropin"``), always in a fresh subprocess so that
``temporal_vit.*`` is imported exactly once per run.

The reference sources are found in ``baseline/_ref`` (installed by tools/install_reference.py, travels to the GPU
box) or, in the build container, in /root/reference.  Nothing here is imported by the product package.
"""
from __future__ import annotations

import glob
import json
import os
import subprocess
import sys
from typing import Dict, List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "shim")

# The fixed synthetic run of BASELINE.json's parity gate ("validation AUC/accuracy parity after a fixed synthetic
# run"): tiny Temporal 3D ViT, dropout 0, fixed init (torch.manual_seed), fixed sample order (shuffle off).
# lr / epochs were chosen so that the reference's own trajectory is stable: a 2 % change of lr moves the final
# val/auc by 0.004 and val/acc by 0.011 (at lr 3e-4 the decision threshold swings by +-0.04 from epoch to epoch and
# no implementation, including the reference on a different BLAS, could be held to a 0.03 accuracy band).
RUN = dict(
    n_trials=4, freq=32, time=64, stride=2,
    train_sessions=64, val_sessions=60, test_sessions=12, trials_per_session=8, signal_prob=0.75, signal_amp=0.35,
    model=dict(model_size="tiny", embed_dim=128, n_heads=2, n_layers=2, dropout=0.0, attention_dropout=0.0,
               drop_path=0.0),
    epochs=12, lr=1e-4, weight_decay=0.01, label_smoothing=0.05, batch_size=16, seed=1234, data_seed=7,
)


def reference_root() -> Optional[str]:
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.exists(os.path.join(cand, "temporal_vit", "training", "train.py")):
            return cand
    return None


def make_synthetic_parquet(out_dir: str, run: dict = RUN) -> Dict[str, str]:
    """Write train/val/test parquet files in the reference's preprocessed layout (data_loader.py:100-128:
    columns session, condition, trial_num, spectrogram = list<list<float>> of shape (freq, time)).

    Sessions alternate WT / FMR1 (label_map {"FMR1": 1}, data_loader.py:94).  Signal: FMR1 trials carry a smooth
    bump of random amplitude in one frequency band on a random subset of the time axis, over N(0,1) noise -- learnable
    but not separable in one epoch, so val AUC moves over the epochs of the run.
    """
    import pandas as pd
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.Generator(np.random.PCG64(run["data_seed"]))
    F, T = run["freq"], run["time"]
    paths = {}
    sid = 0
    for split in ("train", "val", "test"):
        rows: List[dict] = []
        for s in range(run[f"{split}_sessions"]):
            cond = "FMR1" if (s % 2 == 1) else "WT"
            for t in range(run["trials_per_session"]):
                spec = rng.standard_normal((F, T)).astype(np.float32)
                if cond == "FMR1" and rng.random() < run.get("signal_prob", 0.75):
                    t0 = int(rng.integers(0, T - 16))
                    amp = run.get("signal_amp", 0.35) + 0.5 * rng.random()
                    spec[8:16, t0:t0 + 16] += np.float32(amp)
                rows.append({"session": f"s{sid:04d}", "condition": cond, "trial_num": t,
                             "spectrogram": [r.tolist() for r in spec]})
            sid += 1
        path = os.path.join(out_dir, f"{split}.parquet")
        pd.DataFrame(rows).to_parquet(path)
        paths[split] = path
    return paths


_DRIVER = r"""
import json, os, sys
import torch
torch.manual_seed(int(os.environ["TVIT_RUN_SEED"]))
from temporal_vit.training.train import train
from temporal_vit.training.config import TrainConfig
from temporal_vit.data.data_loader import DataLoaderConfig
spec = json.loads(os.environ["TVIT_RUN_SPEC"])
cfg = TrainConfig(
    train_paths=[spec["train"]], val_paths=[spec["val"]], test_paths=[spec["test"]],
    use_preprocessed=True, output_dir=spec["output_dir"], run_name="run", device=spec["device"],
    epochs=spec["epochs"], lr=spec["lr"], weight_decay=spec["weight_decay"], label_smoothing=spec["label_smoothing"],
    loader=DataLoaderConfig(batch_size=spec["batch_size"], shuffle_train=False, num_workers=0),
    n_trials=spec["n_trials"], stride=spec["stride"], **spec["model"])
import temporal_vit.models.model as mm
print("MODEL_MODULE", mm.__file__)
train(cfg)
"""


def run_reference_train(work_dir: str, model: str, device: str, precision: Optional[str] = None, run: dict = RUN,
                        timeout: int = 1800) -> dict:
    """Run train(cfg) in a subprocess.  Returns {"metrics": [per-step dicts], "final": path, "model_module": path}."""
    ref = reference_root()
    if ref is None:
        raise RuntimeError("reference sources not available (neither baseline/_ref nor /root/reference)")
    paths = make_synthetic_parquet(os.path.join(work_dir, "data"), run)
    out_dir = os.path.join(work_dir, f"out_{model}_{precision or 'ref'}")
    spec = dict(paths, output_dir=out_dir, device=device, epochs=run["epochs"], lr=run["lr"],
                weight_decay=run["weight_decay"], label_smoothing=run["label_smoothing"],
                batch_size=run["batch_size"], n_trials=run["n_trials"], stride=run["stride"], model=run["model"])
    env = dict(os.environ)
    pp = [ref] if model == "reference" else [SHIM, ref, ROOT]
    env["PYTHONPATH"] = os.pathsep.join(pp)
    env["TVIT_RUN_SPEC"] = json.dumps(spec)
    env["TVIT_RUN_SEED"] = str(run["seed"])
    if precision:
        env["TVIT_PRECISION"] = precision
    if device == "cpu":
        env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run([sys.executable, "-c", _DRIVER], env=env, cwd=work_dir, capture_output=True, text=True,
                       timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference train() failed ({model}, {device}):\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
    mfiles = glob.glob(os.path.join(out_dir, "run", "metrics", "metrics_*.jsonl"))
    assert len(mfiles) == 1, mfiles
    with open(mfiles[0]) as fh:
        metrics = [json.loads(ln) for ln in fh if ln.strip()]
    module = [ln.split(" ", 1)[1] for ln in r.stdout.splitlines() if ln.startswith("MODEL_MODULE ")][0]
    return {"metrics": metrics, "final": os.path.join(out_dir, "run", "checkpoints", "final.pt"),
            "model_module": module, "stdout": r.stdout}

"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors, host logic, C-ABI
symbol checks, gloo multi-process tests); `-m gpu` tests are the CUDA parity tests proper and run
on a B200.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Load a committed fixture produced by tests/golden/make_golden.py."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    out = {"cfg": {}, "param": {}, "grad": {}, "tap": {}, "mask": {}}
    for k in z.files:
        head, _, tail = k.partition(".")
        if head in out and tail:
            v = z[k]
            if head == "cfg":
                out["cfg"][tail] = v.item()
            elif head == "mask":
                out["mask"][tail] = torch.from_numpy(v.astype(np.float32))
            else:
                out[head][tail] = torch.from_numpy(v)
        else:
            out[k] = z[k]
    return out


def rel_err(a, b):
    """Relative Frobenius error ||a-b|| / ||b|| in float64."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.norm().item()
    if den == 0.0:
        return a.norm().item()
    return (a - b).norm().item() / den

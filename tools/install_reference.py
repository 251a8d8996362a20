#!/usr/bin/env python
"""Install the UNMODIFIED reference package into baseline/_ref (git-ignored, travels with gpurun).

The reference (anthonylu23/neural-vit) ships no setup.py / pyproject, and /root/reference is read-only, so the
documented fallback is used: copy it to a scratch directory under /tmp, add a three-line setup.py THERE, and
``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy>``.  Nothing is written to
/root/reference and no reference source enters the git history.  The installed package is used only as a
checker / baseline:
  * tests that drive the reference's own ``train()`` on the drop-in model (through ``shim/``),
  * ``bench.py --impl reference`` and the ``cpu_baseline`` / ``gpu_eager_baseline`` legs.
Nothing in the product package imports it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")

_SETUP = ('from setuptools import setup, find_packages\n'
          'setup(name="temporal_vit_reference", version="0.0.0",\n'
          '      packages=find_packages(include=["temporal_vit", "temporal_vit.*"]))\n')


def installed() -> bool:
    return os.path.exists(os.path.join(TARGET, "temporal_vit", "models", "model.py"))


def install(force: bool = False, verbose: bool = False) -> bool:
    """Returns True when baseline/_ref holds the reference afterwards."""
    if installed() and not force:
        return True
    if not os.path.isdir(os.path.join(REF_SRC, "temporal_vit")):
        return False                      # e.g. on the GPU box: only a previously installed copy can be used
    tmp = tempfile.mkdtemp(prefix="tvit_ref_")
    try:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, copy, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        with open(os.path.join(copy, "setup.py"), "w") as fh:
            fh.write(_SETUP)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(TARGET, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, copy]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
        return r.returncode == 0 and installed()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv, verbose=True)
    print("baseline/_ref:", "installed" if ok else "unavailable")
    sys.exit(0 if ok else 1)

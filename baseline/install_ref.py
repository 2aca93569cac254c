#!/usr/bin/env python
"""Copy the UNMODIFIED reference into baseline/_ref/ultralytics so that it travels to the GPU box.

    python baseline/install_ref.py

The reference has no setup.py / pyproject.toml (it is an `ultralytics` package directory checked in at the repo
root), so `pip install --target baseline/_ref /root/reference` has nothing to build: pip answers "neither 'setup.py'
nor 'pyproject.toml' found" (recorded in DESIGN.md).  The package imports itself as `ultralytics.*`
(__init__.py:5-10), so the copy is named `ultralytics`; nothing inside is edited -- the five absent third-party
modules (timm, mmcv, mmengine, matplotlib/seaborn, thop) are stubbed in sys.modules by baseline/ref_loader.py at
import time.  baseline/_ref is git-ignored (no reference source enters the history) and not gpurun-ignored.
A MANIFEST (relative path, sha256) is written next to the copy and checked by tests/test_reference_shipped.py.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref", "ultralytics")


def install(force: bool = False) -> str:
    if not os.path.isdir(os.path.join(SRC, "nn", "modules")):
        if os.path.isdir(DST):
            return DST          # GPU box: use what was shipped
        raise RuntimeError(f"{SRC} not present and no shipped copy at {DST}")
    if os.path.isdir(DST):
        if not force and _manifest(SRC) == _read_manifest():
            return DST
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".git"))
    with open(os.path.join(os.path.dirname(DST), "MANIFEST"), "w") as f:
        for rel, h in _manifest(DST):
            f.write(f"{h}  {rel}\n")
    return DST


def _manifest(root):
    out = []
    for d, dirs, files in os.walk(root):
        dirs[:] = sorted(x for x in dirs if x not in ("__pycache__", ".git"))
        for fn in sorted(files):
            if fn.endswith(".pyc"):
                continue
            p = os.path.join(d, fn)
            with open(p, "rb") as fh:
                out.append((os.path.relpath(p, root), hashlib.sha256(fh.read()).hexdigest()))
    return out


def _read_manifest():
    p = os.path.join(os.path.dirname(DST), "MANIFEST")
    if not os.path.isfile(p):
        return None
    out = []
    for line in open(p):
        h, rel = line.rstrip("\n").split("  ", 1)
        out.append((rel, h))
    return out


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))

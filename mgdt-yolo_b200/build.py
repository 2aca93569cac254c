"""Build libmgdt_b200.so in-tree with nvcc for sm_100a (no torch dependency).

    python -m mgdt_yolo_b200.build          # or: from mgdt_yolo_b200.build import build; build()

The .so travels to the GPU box with the repo snapshot (git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmgdt_b200.so")
OBJ = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-DMGDT_WITH_UMMA"]
PER_FILE = {"nms.cu": ["-fmad=false"], "imgproc.cu": ["-fmad=false"], "metrics.cu": ["-fmad=false"]}  # bit-exact IoU / resize arithmetic


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/mgdt_b200.h"]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + fh.read())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, extra=(), out: str | None = None) -> str:
    """`extra` nvcc flags + `out` build a variant library next to the default one (A/B experiments: MGDT_LIB=<out>)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if extra or out:
        return _build_variant(nvcc, list(extra), out or LIB[:-3] + "_variant.so")
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [nvcc, *ARCH, *COMMON, *PER_FILE.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, _sources()))
    r = subprocess.run([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


def _build_variant(nvcc, extra, out):
    objdir = OBJ + "_variant"
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        cmd = [nvcc, *ARCH, *COMMON, *PER_FILE.get(src, []), *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, _sources()))
    r = subprocess.run([nvcc, *ARCH, "-shared", "-o", out, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

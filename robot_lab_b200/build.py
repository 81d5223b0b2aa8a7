"""Build the CUDA extension in-tree: ``python -m robot_lab_b200.build [--force]``.

One translation unit, one shared library (``robot_lab_b200/_lib/libmdpstep.so``), compiled for sm_100a only.
``-fmad=false`` is deliberate (see the header comment of csrc/mdp_step.cu). The built ``.so`` is git-ignored
but travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
SRC = PKG / "csrc" / "mdp_step.cu"
HDR = ROOT / "include" / "rl_mdp_step.h"
OUT = PKG / "_lib" / "libmdpstep.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--split-compile=0", "-lineinfo", "-fmad=false", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", f"-I{ROOT / 'include'}", f"-I{PKG / 'csrc'}",
]


def needs_build() -> bool:
    from . import codegen

    gen = codegen.write()  # rewrites the baked specs only when the task cfgs changed
    if not OUT.exists():
        return True
    newest = max(SRC.stat().st_mtime, HDR.stat().st_mtime, Path(__file__).stat().st_mtime, gen.stat().st_mtime)
    return OUT.stat().st_mtime < newest


def build(force: bool = False, verbose: bool = False) -> Path:
    stale = needs_build()
    if not force and not stale:
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    OUT.parent.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, str(SRC), "-o", str(OUT)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    (OUT.parent / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{log[-4000:]}")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)

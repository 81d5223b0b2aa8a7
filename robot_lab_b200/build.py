"""Build the CUDA extension in-tree: ``python -m robot_lab_b200.build [--force]``.

One shared library (``robot_lab_b200/_lib/libmdpstep.so``), compiled for sm_100a only, from two translation units
that are compiled in parallel and cached as objects: ``csrc/mdp_step.cu`` (the fused step kernels and the C-ABI
around them; minutes to compile because every baked task spec is its own set of kernel instantiations) and
``csrc/scene_kernels.cu`` (the neighbours of the path: actuator models, terrain queries, height-scan casting;
seconds). ``-fmad=false`` is deliberate (see the header comment of csrc/mdp_step.cu). The built ``.so`` is
git-ignored but travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
SOURCES = [CSRC / "mdp_step.cu", CSRC / "scene_kernels.cu"]
HEADERS = [ROOT / "include" / "rl_mdp_step.h", CSRC / "rl_common.cuh"]
OUT = PKG / "_lib" / "libmdpstep.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--split-compile=0", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", f"-I{ROOT / 'include'}", f"-I{CSRC}",
]


def _obj(src: Path) -> Path:
    return OUT.parent / (src.stem + ".o")


def _stale(target: Path, deps: list[Path]) -> bool:
    return (not target.exists()) or target.stat().st_mtime < max(d.stat().st_mtime for d in deps)


def needs_build() -> bool:
    from . import codegen

    gen = codegen.write()  # rewrites the baked specs only when the task cfgs changed
    common = HEADERS + [Path(__file__)]
    deps = {SOURCES[0]: common + [SOURCES[0], gen], SOURCES[1]: common + [SOURCES[1]]}
    return any(_stale(_obj(s), deps[s]) for s in SOURCES) or _stale(OUT, [_obj(s) for s in SOURCES if _obj(s).exists()] + common)


def _compile(nvcc: str, src: Path) -> tuple[int, str]:
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(_obj(src))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    from . import codegen

    gen = codegen.write()
    common = HEADERS + [Path(__file__)]
    deps = {SOURCES[0]: common + [SOURCES[0], gen], SOURCES[1]: common + [SOURCES[1]]}
    todo = [s for s in SOURCES if force or _stale(_obj(s), deps[s])]
    if not todo and not _stale(OUT, [_obj(s) for s in SOURCES]):
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    OUT.parent.mkdir(parents=True, exist_ok=True)
    log = ""
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        for rc, text in pool.map(lambda s: _compile(nvcc, s), todo):
            log += text
            if rc != 0:
                (OUT.parent / "build.log").write_text(log)
                raise RuntimeError(f"nvcc failed:\n{text[-4000:]}")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", *[str(_obj(s)) for s in SOURCES], "-o", str(OUT)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log += " ".join(cmd) + "\n" + res.stdout + res.stderr
    (OUT.parent / "build.log").write_text(log)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout + res.stderr}")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)

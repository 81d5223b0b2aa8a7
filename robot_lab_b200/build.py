"""Build the CUDA extension in-tree: ``python -m robot_lab_b200.build [--force] [--variant NAME]``.

One shared library (``robot_lab_b200/_lib/libmdpstep.so``), compiled for sm_100a only, from two translation units
that are compiled in parallel and cached as objects: ``csrc/mdp_step.cu`` (the fused step kernels and the C-ABI
around them; minutes to compile because every baked task spec is its own set of kernel instantiations) and
``csrc/scene_kernels.cu`` (the neighbours of the path: actuator models, terrain queries, height-scan casting;
seconds). ``-fmad=false`` is deliberate (see the header comment of csrc/mdp_step.cu). The built ``.so`` is
git-ignored but travels to the GPU box with the repo snapshot.

``--variant NAME`` builds an experimental configuration of the step kernels next to the default library
(``_lib/libmdpstep_NAME.so``, loaded with ``RL_MDPSTEP_LIB=...`` for A/B runs and for running the test-suite against
it); the default library and its objects are not touched. Variants: see ``VARIANTS`` below.
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


VARIANTS = {
    "shared_norms": ["-DRL_SHARED_NORMS=1"],               # contact-force norms computed once per env (prepass)
    "shared_ctx": ["-DRL_SHARED_CTX=1"],                   # the three root-frame rotations computed once per env
    "shared": ["-DRL_SHARED_NORMS=1", "-DRL_SHARED_CTX=1"],
    "persistent": ["-DRL_PERSISTENT=1"],                    # a CTA walks several tiles: warm instructions from 16 k envs up
    # same source, no parallel split of the optimiser: --split-compile=0 produced two different codegen "modes" for the
    # same source in round 1; this is the reproducible build to A/B them against ("!" = drop a flag)
    "nosplit": ["!--split-compile=0"],
}


def _obj(src: Path) -> Path:
    return OUT.parent / (src.stem + ".o")


def build_variant(name: str, verbose: bool = False) -> Path:
    """Compile csrc/mdp_step.cu with the variant's defines into its own object / library; scene_kernels.o is shared."""
    from . import codegen

    if name not in VARIANTS:
        raise KeyError(f"unknown variant '{name}'; known: {sorted(VARIANTS)}")
    build()   # the default library (and scene_kernels.o) first
    codegen.write()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    obj = OUT.parent / f"mdp_step.{name}.o"
    out = OUT.parent / f"libmdpstep_{name}.so"
    drop = {f[1:] for f in VARIANTS[name] if f.startswith("!")}
    flags = [f for f in NVCC_FLAGS if f not in drop] + [f for f in VARIANTS[name] if not f.startswith("!")]
    cmd = [nvcc, *flags, "-c", str(SOURCES[0]), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = " ".join(cmd) + "\n" + res.stdout + res.stderr
    (OUT.parent / f"build.{name}.log").write_text(log)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{log[-4000:]}")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", str(obj), str(_obj(SOURCES[1])), "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout + res.stderr}")
    if verbose:
        print(log)
    return out


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
    if "--variant" in sys.argv:
        path = build_variant(sys.argv[sys.argv.index("--variant") + 1], verbose="-v" in sys.argv)
    else:
        path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)

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
SOURCES = [CSRC / "mdp_step.cu", CSRC / "mdp_step_v2.cu", CSRC / "scene_kernels.cu"]
HEADERS = [ROOT / "include" / "rl_mdp_step.h", CSRC / "rl_common.cuh"]
STEP_HEADERS = [CSRC / "mdp_terms.cuh", CSRC / "mdp_ctx.h"]   # shared by the two step-kernel translation units
OUT = PKG / "_lib" / "libmdpstep.so"

NVCC_FLAGS = [
    # no --split-compile: measured in round 2, it makes the generated code differ from build to build (the same source gave
    # step kernels 5 - 7 % apart) for a compile-time gain of seconds; without it two builds are bit-identical
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", f"-I{ROOT / 'include'}", f"-I{CSRC}",
]


# Build variants: experimental configurations of a translation unit, built next to the default library as
# ``_lib/libmdpstep_NAME.so`` and loaded with ``RL_MDPSTEP_LIB=...`` for A/B runs. (The round-1 variants shared_norms /
# shared_ctx / persistent / nosplit were measured in round 2 - profiles/r2_variant_probe.txt - and removed.)
VARIANTS: dict[str, list[str]] = {
    "v2dev": ["-DRL_V2_DEV_ONE=1"],   # cluster kernels for Go2-rough only: seconds to compile while iterating on them
    "stamps": ["-DRL_V2_STAMPS=1"],     # clock stamps per CTA / warp into the debug buffer (tools/v2_timeline.py)
    "unrolled": ["-DRL_V2_UNROLL=1"],   # A/B: term loops of the new kernels fully unrolled against the baked spec
    "few0": ["-DRL_V2_DEV_ONE=1", "-DRL_V2_FEW_UNROLL=0"],   # A/B: the short loops rolled (Go2-rough only)
    "stamps_few0": ["-DRL_V2_DEV_ONE=1", "-DRL_V2_STAMPS=1", "-DRL_V2_FEW_UNROLL=0"],
    # A/B: register budget of the new kernels (Go2-rough only) - resident CTAs per SM at 8 / 16 warps per CTA
    "occ6": ["-DRL_V2_DEV_ONE=1", "-DRL_V2_PRE_MINB8=6", "-DRL_V2_PRE_MINB16=3"],
    "occ8": ["-DRL_V2_DEV_ONE=1", "-DRL_V2_PRE_MINB8=8", "-DRL_V2_PRE_MINB16=4", "-DRL_V2_POST_MINB8=8", "-DRL_V2_POST_MINB16=4"],
}


def _obj(src: Path) -> Path:
    return OUT.parent / (src.stem + ".o")


def build_variant(name: str, verbose: bool = False) -> Path:
    """Compile the step-kernel translation units with the variant's defines into their own objects / library."""
    from . import codegen

    if name not in VARIANTS:
        raise KeyError(f"unknown variant '{name}'; known: {sorted(VARIANTS)}")
    build()   # the default library (and the objects the variant shares with it) first
    codegen.write()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    out = OUT.parent / f"libmdpstep_{name}.so"
    flags = NVCC_FLAGS + VARIANTS[name]
    objs = []
    for src in SOURCES:
        if src.name != "mdp_step_v2.cu":   # the variants so far only touch the cluster kernels
            objs.append(_obj(src))
            continue
        obj = OUT.parent / f"{src.stem}.{name}.o"
        cmd = [nvcc, *flags, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = " ".join(cmd) + "\n" + res.stdout + res.stderr
        (OUT.parent / f"build.{name}.log").write_text(log)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{log[-4000:]}")
        if verbose:
            print(log)
        objs.append(obj)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", *[str(o) for o in objs], "-o", str(out)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout + res.stderr}")
    return out


def _stale(target: Path, deps: list[Path]) -> bool:
    return (not target.exists()) or target.stat().st_mtime < max(d.stat().st_mtime for d in deps)


def needs_build() -> bool:
    from . import codegen

    gen = codegen.write()  # rewrites the baked specs only when the task cfgs changed
    common = HEADERS + [Path(__file__)]
    deps = _deps(common, gen)
    return any(_stale(_obj(s), deps[s]) for s in SOURCES) or _stale(OUT, [_obj(s) for s in SOURCES if _obj(s).exists()] + common)


def _deps(common: list[Path], gen: Path) -> dict[Path, list[Path]]:
    return {src: common + [src] + (STEP_HEADERS + [gen] if src.name.startswith("mdp_step") else []) for src in SOURCES}


def _compile(nvcc: str, src: Path) -> tuple[int, str]:
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(_obj(src))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    from . import codegen

    gen = codegen.write()
    common = HEADERS + [Path(__file__)]
    deps = _deps(common, gen)
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

"""What the compiled step kernels actually contain (cuobjdump -sass of the in-tree library, no GPU needed): the TMA
bulk copies and async copies the load phase is built on, the release-ticket tails, the named barriers of the
two-tile configuration, sm_100a as the only target, and (next to) no local-memory traffic in the production kernels
(16 warps per tile)."""

import os
import re
import shutil
import subprocess

import pytest

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def kernels(native_lib):
    from robot_lab_b200 import _native as nat

    if not shutil.which(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    lib_path = os.environ.get("RL_MDPSTEP_LIB", str(nat.LIB_PATH))   # the library _native.load() picks
    out = subprocess.run([CUOBJDUMP, "-sass", lib_path], capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in out and not re.search(r"arch = sm_(?!100a)", out), "the library targets sm_100a only"
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None and ";" in line:
            funcs[cur].append(line)
    return funcs


def _step_kernels(funcs, nw=16, dbg=0, tiles=1, static=True):
    pat = re.compile(r"mdp_step_kernelINS_(\w+?)E?Li%dELi0ELb%dELi%dEEEv" % (nw, dbg, tiles))
    return {k: v for k, v in funcs.items() if (m := pat.search(k)) and (("StaticPolicy" in m.group(1)) == static)}


def test_load_phase_uses_tma_bulk_copies_and_async_copies(kernels):
    ks = _step_kernels(kernels)
    assert len(ks) >= 5, "one production instantiation per baked task"
    for name, body in ks.items():
        text = "\n".join(body)
        assert "UBLKCP" in text, f"{name}: no TMA bulk copy (cp.async.bulk)"
        assert "LDGSTS" in text, f"{name}: no cp.async"
        assert "SYNCS" in text, f"{name}: no mbarrier"


def test_last_cta_tickets_are_release_atomics_not_sc_fences(kernels):
    for name, body in _step_kernels(kernels).items():
        text = "\n".join(body)
        assert len(re.findall(r"ATOM\.E\.ADD\.STRONG\.GPU", text)) >= 2, name          # early arrivals (pre- / post-reset)
        assert len(re.findall(r"MEMBAR\.ALL\.GPU", text)) >= 2, name                   # their release fences
        assert len(re.findall(r"MEMBAR\.SC\.GPU", text)) <= 2, name                    # only the two acquiring tails


def test_two_tile_kernels_use_named_barriers_of_512_threads(kernels):
    ks = _step_kernels(kernels, tiles=2)
    assert len(ks) >= 5
    for name, body in ks.items():
        text = "\n".join(body)
        assert re.search(r"BAR\.SYNC\.DEFER_BLOCKING R\d+, 0x200", text), name
        assert re.search(r"BAR\.RED\.OR\.DEFER_BLOCKING R\d+, 0x200", text), name
        assert not re.search(r"BAR\.SYNC\.DEFER_BLOCKING 0x0", text), f"{name}: a CTA-wide barrier would dead-lock the tiles"


def test_production_kernels_keep_their_state_in_registers(kernels):
    """Local-memory instructions only as the few callee-save slots around the shared (noinline) helpers: well under 1 %
    of a kernel (the Go2-rough kernels of the round-1 build have none; ptxas -v per build: _lib/build.log)."""
    for tiles in (1, 2):
        for name, body in _step_kernels(kernels, tiles=tiles).items():
            local = sum(1 for line in body if re.search(r"\b(STL|LDL)\b", line))
            assert local * 100 <= len(body), f"{name}: {local} local-memory instructions of {len(body)}"

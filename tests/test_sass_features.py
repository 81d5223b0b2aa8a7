"""What the compiled step kernels actually contain (cuobjdump -sass of the in-tree library, no GPU needed): the 2-D TMA
tensor-map copies (UTMALDG), the bulk copies, the cluster barriers and DSMEM addressing of the cluster kernels
(csrc/mdp_step_v2.cu); the bulk / async copies and release-ticket tails of the general kernel (csrc/mdp_step.cu);
sm_100a as the only target; (next to) no local-memory traffic in the production kernels."""

import os
import re
import shutil
import subprocess

import pytest

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
N_BAKED = 7   # tasks with a build-time specialised spec (csrc/generated/baked_specs.cuh)


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


def _step_kernels(funcs, nw=16, dbg=0, static=True):
    """Instantiations of the general kernel mdp_step_kernel<Policy, NW, MODE = 0, DBG>."""
    pat = re.compile(r"mdp_step_kernelINS_(\w+?)E?Li%dELi0ELb%dEEEv" % (nw, dbg))
    return {k: v for k, v in funcs.items() if (m := pat.search(k)) and (("StaticPolicy" in m.group(1)) == static)}


def _cluster_kernels(funcs, kind, c=None, g=None):
    """Instantiations of v2_pre_kernel / v2_post_kernel <Baked, C, G, NW>."""
    pat = re.compile(r"v2_%s_kernelIN5baked\w+?ELi(\d+)ELi(\d+)ELi\d+EEEv" % kind)
    return {k: v for k, v in funcs.items()
            if (m := pat.search(k)) and (c is None or int(m.group(1)) == c) and (g is None or int(m.group(2)) == g)}


def test_cluster_kernels_stage_fields_with_tensor_map_copies(kernels):
    """Every cluster kernel: SoA fields by cp.async.bulk.tensor (UTMALDG) issued from a compact loop, the mbarrier they
    complete on, and NO per-thread async-copy loops (LDGSTS) - the load prologue of the general kernel is gone."""
    for kind in ("pre", "post"):
        ks = _cluster_kernels(kernels, kind)
        assert len(ks) == 5 * N_BAKED, f"{kind}: {len(ks)} kernels, want 5 configurations x {N_BAKED} tasks"
        for name, body in ks.items():
            text = "\n".join(body)
            assert "UTMALDG.2D" in text, f"{name}: no 2-D tensor-map copy"
            assert "SYNCS" in text, f"{name}: no mbarrier"
            assert "LDGSTS" not in text, f"{name}: per-thread async copies are back"
    for name, body in _cluster_kernels(kernels, "pre").items():
        assert "UBLKCP" not in "\n".join(body), f"{name}: the contact-force rows are streamed from global memory, not staged"


def test_cluster_kernels_use_cluster_barriers_and_dsmem(kernels):
    for kind in ("pre", "post"):
        for (c, g) in ((4, 4), (2, 2)):
            ks = _cluster_kernels(kernels, kind, c, g)
            assert len(ks) == N_BAKED
            for name, body in ks.items():
                text = "\n".join(body)
                assert "UCGABAR_ARV" in text and "UCGABAR_WAIT" in text, f"{name}: no cluster barrier"
        one = _cluster_kernels(kernels, kind, 1, 1)
        assert len(one) == 3 * N_BAKED   # 16, 8 and 4 warps per tile
        for name, body in one.items():
            assert "UCGABAR" not in "\n".join(body), f"{name}: a one-CTA configuration needs no cluster barrier"


def test_new_kernels_are_smaller_than_the_general_kernel(kernels):
    """No load prologue, no store phase, no per-warp protocols: a new kernel is under half of the general kernel of the
    same task (5.3 - 5.6 k instructions in round 1). Code size is time here: every SM executes each instruction once."""
    for kind, bound in (("pre", 5200), ("post", 5200)):
        for name, body in _cluster_kernels(kernels, kind).items():
            assert len(body) < bound, f"{name}: {len(body)} instructions"


def test_general_kernel_load_phase_uses_bulk_and_async_copies(kernels):
    ks = _step_kernels(kernels)
    assert len(ks) == N_BAKED, "one production instantiation per baked task"
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
    # the pre-reset kernel orders the reset ids by a decoupled look-back (plain relaxed loads / stores of one status word per
    # tile: no atomic, no fence, no launch-wide tail); the post-reset kernel's logging reduction keeps one release ticket per
    # LOG part and an acquiring LOAD by the thread that drew the last one (no SC fence)
    for name, body in _cluster_kernels(kernels, "pre").items():
        text = "\n".join(body)
        assert "ATOM" not in text, name
        assert "MEMBAR.SC" not in text, name
    for name, body in _cluster_kernels(kernels, "post").items():
        text = "\n".join(body)
        assert len(re.findall(r"ATOM\.E\.ADD\.STRONG\.GPU", text)) >= 1, name
        assert "MEMBAR.SC" not in text, name


def test_production_kernels_keep_their_state_in_registers(kernels):
    """Local-memory instructions only as the few callee-save slots around the shared (noinline) helpers: well under 1 %
    of a kernel."""
    groups = [_step_kernels(kernels), _cluster_kernels(kernels, "pre"), _cluster_kernels(kernels, "post")]
    for ks in groups:
        assert ks
        for name, body in ks.items():
            local = sum(1 for line in body if re.search(r"\b(STL|LDL)\b", line))
            assert local * 100 <= len(body), f"{name}: {local} local-memory instructions of {len(body)}"

"""Golden fixtures (tests/golden/*.npz, produced by the reference's own functions - see make_golden.py) pin the
oracle wherever it runs, and on a GPU pin the CUDA kernels to the reference directly."""

from pathlib import Path

import numpy as np
import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200.synthetic import make_state

GOLDEN = Path(__file__).resolve().parent / "golden"
KEYS = ["a1_flat", "go2_flat", "go2_rough", "g1_rough", "g1_rough_37", "go2_catalogue"]


def _load(key):
    z = np.load(GOLDEN / f"{key}.npz")
    st = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in/")}
    rew = {k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("reward/")}
    cmd = {k[8:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("command/")}
    obs = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("obs/")}
    return st, rew, cmd, obs


def _spec(key):
    return H.make_catalogue_spec() if key == "go2_catalogue" else H.make_spec(key)


@pytest.mark.parametrize("key", KEYS)
def test_oracle_reproduces_reference_outputs(key):
    cfg, spec = _spec(key)
    st, rew, cmd, _ = _load(key)
    d = port.Derived(st, spec)
    by_name = {t.name: t for t in spec.rewards}
    assert rew and set(rew) <= set(by_name)
    for name, want in rew.items():
        torch.testing.assert_close(port.reward_term(by_name[name], st, spec, d), want, rtol=1e-6, atol=1e-6, msg=name)
    if cmd:
        got = port.compute_command(spec, st, {"cmd_uniforms": st["cmd_uniforms"]})
        for k, v in cmd.items():
            assert torch.equal(got[k], v), k


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_cuda_terms_reproduce_reference_outputs(native_lib, key):
    """CUDA vs the reference's own numbers (no oracle in between)."""
    from robot_lab_b200.engine import MdpStepEngine

    cfg, spec = _spec(key)
    st, rew, cmd, _ = _load(key)
    n = st["root_quat_w"].shape[0]
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n)
    b.load_logical(st)
    by_name = {t.name: t for t in spec.rewards}
    for name, want in rew.items():
        got = eng.term_eval(by_name[name], b).cpu()
        torch.testing.assert_close(got, want, rtol=H.RTOL, atol=H.ATOL, msg=name)
    if cmd:
        from robot_lab_b200 import _native as nat

        eng.step(b, phases=nat.PHASE_COMMAND)
        torch.cuda.synchronize()
        for k, v in cmd.items():
            g = b.logical(k).cpu().contiguous()
            if v.dtype == torch.bool:
                assert torch.equal(g, v), k
            else:
                torch.testing.assert_close(g, v, rtol=H.RTOL, atol=H.ATOL, msg=k)
    eng.close()


@pytest.mark.parametrize("key", ["go2_rough", "g1_rough"])
def test_oracle_reset_event_reproduces_reference_output(key):
    """oracle.reset_scene_state vs the committed output of the reference's reset_root_state_uniform."""
    import numpy as np

    from robot_lab_b200.cfg import ResetStateCfg

    z = np.load(GOLDEN / f"reset_state_{key}.npz")
    cfg, spec = H.make_spec(key)
    n = z["uniforms"].shape[1]
    st = make_state(spec, n, seed=20260922)
    ids = torch.from_numpy(z["ids"])
    got = port.reset_scene_state(spec, st, ids, ResetStateCfg.go2_rough(), torch.from_numpy(z["env_origins"]),
                                 torch.from_numpy(z["uniforms"]))
    for k in ("root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w"):
        torch.testing.assert_close(got[k][ids.long()], torch.from_numpy(z[f"out/{k}"]), rtol=0, atol=0, msg=k)

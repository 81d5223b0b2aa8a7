"""rl_reset_scene_state (reset_root_state_uniform, V/mdp/events.py:205-271, + reset_joints_by_scale
[IL]) through the C-ABI: against the oracle, and against the committed outputs of the reference function itself."""
from pathlib import Path

import numpy as np
import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200.cfg import ResetStateCfg
from robot_lab_b200.engine import MdpStepEngine
from robot_lab_b200.synthetic import make_state

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
KEYS = ("root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w", "joint_pos", "joint_vel")


def _check(b, want, atol=2e-6):
    for k in KEYS:
        torch.testing.assert_close(b.logical(k).cpu().contiguous(), want[k], rtol=H.RTOL, atol=atol, msg=k)


@pytest.mark.parametrize("key,n,layout", [("go2_rough", 777, "soa"), ("g1_rough", 130, "aos")])
def test_id_list_and_mask_forms_match_oracle(native_lib, key, n, layout):
    cfg, spec = H.make_spec(key)
    st = make_state(spec, n)
    g = torch.Generator().manual_seed(5)
    u = torch.rand(12 + 2 * spec.J, n, generator=g)
    org = torch.randn(n, 3, generator=g) * 30.0
    rc = ResetStateCfg(pose_range={"x": (-0.5, 0.5), "y": (-0.5, 0.5), "z": (0.0, 0.2), "roll": (-3.14, 3.14),
                                   "pitch": (-3.14, 3.14), "yaw": (-3.14, 3.14)},
                       joint_position_range=(0.5, 1.5), joint_velocity_range=(-1.0, 1.0))
    eng = MdpStepEngine(spec, "cuda:0")
    # (a) id list
    b = eng.new_buffers(n, layout=layout)
    b.load_logical(st)
    ids = torch.randperm(n, generator=g)[: n // 4].sort().values.int()
    eng.reset_scene_state(b, rc, org.cuda(), env_ids=ids.cuda(), n_env_ids=torch.tensor([len(ids)], dtype=torch.int32).cuda(),
                          uniforms=u.cuda())
    torch.cuda.synchronize()
    _check(b, port.reset_scene_state(spec, st, ids, rc, org, u))
    # (b) the envs flagged done by the pre-reset launch
    b2 = eng.new_buffers(n, layout=layout)
    b2.load_logical(st)
    eng.step_pre_reset(b2)
    eng.reset_scene_state(b2, rc, org.cuda(), uniforms=u.cuda())
    torch.cuda.synchronize()
    done = (b2.terminated.bool() | b2.truncated.bool()).cpu()
    assert 0 < int(done.sum()) < n
    _check(b2, port.reset_scene_state(spec, st, done.nonzero().flatten(), rc, org, u))
    eng.close()


@pytest.mark.parametrize("key", ["go2_rough", "g1_rough"])
def test_matches_the_reference_functions_own_output(native_lib, key):
    z = np.load(GOLDEN / f"reset_state_{key}.npz")
    cfg, spec = H.make_spec(key)
    n = z["uniforms"].shape[1]
    st = make_state(spec, n, seed=20260922)
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n)
    b.load_logical(st)
    ids = torch.from_numpy(z["ids"])
    eng.reset_scene_state(b, ResetStateCfg.go2_rough(), torch.from_numpy(z["env_origins"]).cuda(), env_ids=ids.cuda(),
                          n_env_ids=torch.tensor([len(ids)], dtype=torch.int32).cuda(),
                          uniforms=torch.from_numpy(z["uniforms"]).cuda())
    torch.cuda.synchronize()
    for k in ("root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w"):
        torch.testing.assert_close(b.logical(k).cpu()[ids.long()], torch.from_numpy(z[f"out/{k}"]), rtol=H.RTOL, atol=2e-6, msg=k)
    eng.close()


def test_philox_mode_is_reproducible_and_in_range(native_lib):
    cfg, spec = H.make_spec("go2_rough")
    n = 512
    st = make_state(spec, n)
    eng = MdpStepEngine(spec, "cuda:0")
    rc = ResetStateCfg.go2_rough()
    ids = torch.arange(n, dtype=torch.int32).cuda()
    cnt = torch.tensor([n], dtype=torch.int32).cuda()
    outs = []
    for seed in (7, 7, 8):
        b = eng.new_buffers(n)
        b.load_logical(st)
        eng.reset_scene_state(b, rc, None, env_ids=ids, n_env_ids=cnt, seed=seed)
        torch.cuda.synchronize()
        outs.append({k: b.logical(k).clone() for k in KEYS})
    assert all(torch.equal(outs[0][k], outs[1][k]) for k in KEYS)
    assert not torch.equal(outs[0]["root_pos_w"], outs[2]["root_pos_w"])
    p = outs[0]["root_pos_w"].cpu()
    z0 = spec.layout.asset.init_root_height
    assert p[:, 0].abs().max() <= 0.5 and p[:, 1].abs().max() <= 0.5 and (p[:, 2] >= z0).all() and (p[:, 2] <= z0 + 0.2 + 1e-6).all()
    torch.testing.assert_close(outs[0]["root_quat_w"].norm(dim=-1).cpu(), torch.ones(n), rtol=1e-5, atol=1e-5)
    assert outs[0]["root_lin_vel_w"].abs().max() <= 0.5
    assert 0.2 < p[:, 0].std() < 0.4     # U(-0.5, 0.5): std 0.289
    eng.close()

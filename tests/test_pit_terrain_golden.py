"""The "pits" code paths of the reference (V/mdp/utils.py:44-127, V/mdp/commands.py:61-85, V/mdp/events.py:232-244):
committed outputs of the reference's own functions on a terrain WITH a "pits" sub-terrain
(tests/golden/pit_terrain_*.npz, tests/golden/make_golden.py::pits) pin the oracle everywhere and, on a GPU, the
CUDA kernels rl_is_robot_on_terrain / rl_command_pit_restrict / rl_reset_scene_state directly."""
from pathlib import Path

import numpy as np
import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200 import terrain as terrain_host
from robot_lab_b200.cfg import ResetStateCfg, TerrainCfg

GOLDEN = Path(__file__).resolve().parent / "golden"
TER = TerrainCfg(sub_terrains=("pyramid_stairs", "pits", "boxes", "random_rough", "hf_pyramid_slope"),
                 proportions=(0.2, 0.15, 0.25, 0.3, 0.1))
CMD_KEYS = ("command", "heading_target", "time_left", "is_heading_env", "is_standing_env", "metric_error_vel_xy",
            "metric_error_vel_yaw")
ROOT_KEYS = ("root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w")


def _load(key):
    z = np.load(GOLDEN / f"pit_terrain_{key}.npz")
    t = lambda k: torch.from_numpy(z[k])  # noqa: E731
    st = {k[3:]: t(k) for k in z.files if k.startswith("in/")}
    return z, t, st


@pytest.mark.parametrize("key", ["go2_rough", "g1_rough"])
def test_oracle_reproduces_reference_pit_outputs(key):
    z, t, st = _load(key)
    cfg, spec = H.make_spec(key)
    origins, types, was = t("terrain_origins"), t("terrain_types"), t("was_on_pit")
    for name in ("pits", "boxes"):
        rng = terrain_host.terrain_column_range(TER, name)
        assert torch.equal(port.is_robot_on_terrain(st["root_pos_w"], origins, rng), t(f"on_terrain/{name}"))
        assert torch.equal(terrain_host.is_env_assigned_to_terrain(TER, types, name).bool(), t(f"assigned/{name}"))
    u = st["cmd_uniforms"]
    got = port.compute_command(spec, st, {"cmd_uniforms": u})
    on = port.is_robot_on_terrain(st["root_pos_w"], origins, terrain_host.terrain_column_range(TER, "pits"))
    got.update(port.command_pit_restrict(spec, {**st, **got}, on, was, u))
    for k in CMD_KEYS + ("was_on_pit",):
        assert torch.equal(got[k], t(f"command/{k}")), k
    ids = t("reset/ids")
    st_full = {**st, "joint_pos": torch.zeros(len(was), spec.J), "joint_vel": torch.zeros(len(was), spec.J)}
    rs = port.reset_scene_state(spec, st_full, ids, ResetStateCfg.go2_rough(), t("reset/env_origins"), t("reset/uniforms"),
                                assigned_to_pits=t("assigned/pits"))
    for k in ROOT_KEYS:
        torch.testing.assert_close(rs[k][ids.long()], t(f"reset/out/{k}"), rtol=0, atol=0, msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("key,layout", [("go2_rough", "soa"), ("g1_rough", "aos")])
def test_cuda_reproduces_reference_pit_outputs(native_lib, key, layout):
    """CUDA vs the reference's own numbers: terrain lookup and flags bit-exact, floats within 1e-5 relative."""
    from robot_lab_b200 import _native as nat
    from robot_lab_b200.engine import MdpStepEngine

    z, t, st = _load(key)
    cfg, spec = H.make_spec(key)
    n = st["root_pos_w"].shape[0]
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n, layout=layout)
    b.load_logical(st)
    for name in ("pits", "boxes"):
        grid = terrain_host.TerrainGridBuffers.create(TER, name, "cuda:0", origins=t("terrain_origins"))
        got = eng.is_robot_on_terrain(b, grid).cpu().bool()
        assert torch.equal(got, t(f"on_terrain/{name}")), name
    grid = terrain_host.TerrainGridBuffers.create(TER, "pits", "cuda:0", origins=t("terrain_origins"))
    was = t("was_on_pit").to(torch.uint8).cuda()
    eng.step(b, phases=nat.PHASE_COMMAND)          # CommandTerm.compute up to the parent's _update_command
    eng.command_pit_restrict(b, grid, was)         # ... and its terrain-aware tail
    torch.cuda.synchronize()
    assert torch.equal(was.cpu().bool(), t("command/was_on_pit"))
    for k in CMD_KEYS:
        g, v = b.logical(k).cpu().contiguous(), t(f"command/{k}")
        if v.dtype == torch.bool:
            assert torch.equal(g, v), k
        else:
            torch.testing.assert_close(g, v, rtol=H.RTOL, atol=H.ATOL, msg=k)
    # the envs on pits carry exactly the restricted command
    on = t("command/was_on_pit")
    c = b.logical("command").cpu()
    assert ((c[on, 0] >= 0.3) & (c[on, 0] <= 0.6)).all() and (c[on, 1:] == 0).all()
    # reset event with the pit branch
    ids = t("reset/ids")
    eng.reset_scene_state(b, ResetStateCfg.go2_rough(), t("reset/env_origins").cuda(), env_ids=ids.cuda(),
                          n_env_ids=torch.tensor([len(ids)], dtype=torch.int32).cuda(), uniforms=t("reset/uniforms").cuda(),
                          assigned_to_pits=t("assigned/pits").to(torch.uint8).cuda())
    torch.cuda.synchronize()
    for k in ROOT_KEYS:
        torch.testing.assert_close(b.logical(k).cpu()[ids.long()], t(f"reset/out/{k}"), rtol=H.RTOL, atol=2e-6, msg=k)
    pit_ids = ids.long()[t("assigned/pits")[ids.long()]]
    assert len(pit_ids) > 0 and (b.logical("root_lin_vel_w").cpu()[pit_ids] == 0).all()
    eng.close()

"""SURVEY.md 8(f) rows 3 and 4 through the C-ABI on a GPU: rl_actuator_step, rl_is_robot_on_terrain,
rl_command_pit_restrict, rl_height_scan_cast against the oracle on the same seeded inputs (both tensor layouts),
plus size-independent properties at the benchmark's env count."""
import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200 import _native as nat, terrain as terrain_host
from robot_lab_b200.cfg import RayCasterCfg, TerrainCfg
from robot_lab_b200.engine import MdpStepEngine
from robot_lab_b200.synthetic import make_state

pytestmark = pytest.mark.gpu
PIT_TERRAIN = TerrainCfg(sub_terrains=("pyramid_stairs", "pits", "boxes", "random_rough", "hf_pyramid_slope"),
                         proportions=(0.2, 0.15, 0.25, 0.3, 0.1))


@pytest.mark.parametrize("key,n,layout", [("go2_rough", 4096, "soa"), ("a1_flat", 64, "aos"), ("g1_rough", 1000, "soa"),
                                          ("g1_rough_37", 333, "aos")])
def test_actuator_step_matches_oracle(native_lib, key, n, layout):
    cfg, spec = H.make_spec(key)
    st = make_state(spec, n, seed=31)
    g = torch.Generator().manual_seed(32)
    # targets around the joint position so that the PD law lands inside, at and beyond the limits; joint speeds up to
    # and beyond the no-load speed so that every branch of the torque-speed curve is taken
    tab = spec.layout.asset.actuator_table()
    vlim = torch.tensor([v if v > 0 else 30.0 for v in tab["velocity_limit"]])
    st["joint_vel"] = (torch.rand(n, spec.J, generator=g) * 2.8 - 1.4) * vlim
    target = st["joint_pos"] + torch.randn(n, spec.J, generator=g) * 0.6
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n, layout=layout)
    b.load_logical(st)
    b._to_device("joint_target", target)
    comp = torch.zeros((spec.J, n) if layout == "soa" else (n, spec.J), device="cuda")
    eng.actuator_step(b, computed_torque=comp)
    torch.cuda.synchronize()
    want_c, want_a = port.actuator_step(tab, target, st["joint_pos"], st["joint_vel"])
    got_c = (comp.t() if layout == "soa" else comp).cpu()
    torch.testing.assert_close(got_c, want_c, rtol=H.RTOL, atol=H.ATOL)
    torch.testing.assert_close(b.logical("applied_torque").cpu().contiguous(), want_a, rtol=H.RTOL, atol=H.ATOL)
    if "dc_motor" in tab["kind"]:
        lim = torch.tensor(tab["effort_limit"])
        a = b.logical("applied_torque").cpu()
        assert (a.abs() <= lim + 1e-6).all()
        assert ((a - want_c).abs() > 1e-3).float().mean() > 0.05    # the clip is exercised ...
        assert ((a - want_c).abs() < 1e-6).float().mean() > 0.05    # ... and so is the linear range
        # a motor spinning faster than its no-load speed cannot push further in that direction
        fast = st["joint_vel"] > vlim * 1.01
        assert fast.any() and (a[fast] <= 1e-6).all()
    eng.close()


def test_actuator_step_rejects_bad_arguments(native_lib):
    import ctypes as C

    cfg, spec = H.make_spec("go2_rough")
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(32)
    bad = nat.RlActuatorCfg()
    bad.num_joints = spec.J + 1
    st, tgt = b.state_view(), b.field("joint_target")
    rc = eng.lib.rl_actuator_step(eng._ctx, 32, C.byref(bad), C.byref(tgt), None, None, C.byref(st), None, 0)
    assert rc == -1 and b"joints" in eng.lib.rl_last_error()
    ok = eng.actuator_cfg()
    assert eng.lib.rl_actuator_step(eng._ctx, 0, C.byref(ok), C.byref(tgt), None, None, C.byref(st), None, 0) == 0  # empty
    eng.close()


@pytest.mark.parametrize("n,layout", [(4096, "soa"), (777, "aos")])
def test_terrain_lookup_and_pit_restrict_match_oracle(native_lib, n, layout):
    cfg, spec = H.make_spec("go2_rough")
    st = make_state(spec, n, seed=41)
    g = torch.Generator().manual_seed(42)
    origins = terrain_host.grid_origins(PIT_TERRAIN)
    rng = terrain_host.terrain_column_range(PIT_TERRAIN, "pits")
    grid = terrain_host.TerrainGridBuffers.create(PIT_TERRAIN, "pits", "cuda:0")
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n, layout=layout)
    was = torch.rand(n, generator=g) < 0.3
    was_dev = was.to(torch.uint8).cuda()
    u = st["cmd_uniforms"]
    for it in range(3):   # robots move between the steps: enter, stay on, leave the pits
        st["root_pos_w"] = torch.stack([(torch.rand(n, generator=g) - 0.5) * 110.0, (torch.rand(n, generator=g) - 0.5) * 190.0,
                                        torch.rand(n, generator=g)], dim=1)
        b.load_logical(st)
        on = port.is_robot_on_terrain(st["root_pos_w"], origins, rng)
        assert torch.equal(eng.is_robot_on_terrain(b, grid).cpu().bool(), on)
        eng.step(b, phases=nat.PHASE_COMMAND)
        eng.command_pit_restrict(b, grid, was_dev)
        torch.cuda.synchronize()
        want = port.compute_command(spec, st, {"cmd_uniforms": u})
        want.update(port.command_pit_restrict(spec, {**st, **want}, on, was, u))
        got = H.gpu_step_outputs(b)
        H.compare_outputs(got, want, keys=("command", "heading_target", "time_left", "is_heading_env", "is_standing_env",
                                          "metric_error_vel_xy", "metric_error_vel_yaw"))
        assert torch.equal(was_dev.cpu().bool(), on)
        assert int((was & ~on).sum()) > 10 and int(on.sum()) > 10
        was = on
        for k in ("command", "heading_target", "time_left", "is_heading_env", "is_standing_env", "metric_error_vel_xy",
                  "metric_error_vel_yaw"):
            st[k] = want[k]
    eng.close()


def test_pit_restrict_philox_stream_matches_oracle_port(native_lib):
    """Production randomness: the left-pit resample draws from Philox stream 5 (blocks 0, 1)."""
    from oracle import philox

    cfg, spec = H.make_spec("go2_rough")
    n = 2048
    st = make_state(spec, n, seed=43)
    st["root_pos_w"][:, 1] = 0.0     # column 9 or 10: nobody is on the pits (columns 4..6) ...
    grid = terrain_host.TerrainGridBuffers.create(PIT_TERRAIN, "pits", "cuda:0")
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n)
    b.load_logical(st)
    was_dev = torch.ones(n, dtype=torch.uint8, device="cuda")   # ... and everybody just left one
    eng.command_pit_restrict(b, grid, was_dev, seed=99, step=7, env_id_offset=4096, use_random_inputs=False)
    torch.cuda.synchronize()
    u = philox.command_uniforms(n, 99, 7, 4096, 5)
    on = torch.zeros(n, dtype=torch.bool)
    want = port.command_pit_restrict(spec, st, on, torch.ones(n, dtype=torch.bool), u)
    got = H.gpu_step_outputs(b)
    H.compare_outputs(got, want, keys=("command", "heading_target", "is_heading_env", "is_standing_env"))
    assert not was_dev.any()
    eng.close()


def _rough_height_field(nx, ny, seed):
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(nx // 8 + 2, ny // 8 + 2, generator=g) * 0.6
    up = torch.nn.functional.interpolate(base[None, None], size=(nx, ny), mode="bilinear", align_corners=True)[0, 0]
    steps = torch.round(torch.rand(nx, ny, generator=g) * 4.0) * 0.005   # height-field quantisation noise [IL]
    return (up + steps).contiguous()


@pytest.mark.parametrize("key,n,layout", [("go2_rough", 4096, "soa"), ("g1_rough", 300, "aos")])
def test_height_scan_cast_matches_oracle(native_lib, key, n, layout):
    cfg, spec = H.make_spec(key)
    assert spec.R == 187
    st = make_state(spec, n, seed=51)
    g = torch.Generator().manual_seed(52)
    nx, ny, hs = 400, 640, 0.1
    x0, y0 = -20.0, -32.0
    heights = _rough_height_field(nx, ny, 53)
    st["root_pos_w"] = torch.stack([(torch.rand(n, generator=g) - 0.5) * 41.0, (torch.rand(n, generator=g) - 0.5) * 65.0,
                                    torch.rand(n, generator=g) + 0.3], dim=1)   # a few robots beyond the border
    starts = terrain_host.grid_pattern_ray_starts(RayCasterCfg())
    hf = terrain_host.HeightFieldBuffers(heights.cuda(), x0, y0, hs, starts.cuda())
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n, layout=layout)
    b.load_logical(st)
    eng.height_scan_cast(b, hf)
    torch.cuda.synchronize()
    want_z, want_s = port.height_scan_cast(heights, x0, y0, hs, starts, st["root_pos_w"], st["root_quat_w"])
    got_z = b.logical("ray_hits_z").cpu().contiguous()
    # sin / cos / atan2 of the yaw differ by an ulp between libm and CUDA: the ray's xy moves by ~1e-6 m, which is
    # ~1e-5 m of height across the steepest cells of this field - and may flip a ray that grazes the border
    assert int((torch.isinf(got_z) != torch.isinf(want_z)).sum()) <= 2 and torch.isinf(want_z).any()
    fin = torch.isfinite(want_z) & torch.isfinite(got_z)
    torch.testing.assert_close(got_z[fin], want_z[fin], rtol=0, atol=2e-5)
    assert (got_z[fin] - want_z[fin]).abs().mean() < 1e-6
    assert torch.equal(b.logical("ray_sensor_pos_z").cpu(), want_s)
    # ... and the observation built from it equals the oracle's height_scan term
    st["ray_hits_z"], st["ray_sensor_pos_z"] = want_z, want_s
    eng.step(b, phases=nat.PHASE_OBS)
    torch.cuda.synchronize()
    ref = port.compute_obs_group(spec, 1, st, H.rnd_inputs(st))
    same = (torch.isinf(got_z) == torch.isinf(want_z)).all(dim=1)
    torch.testing.assert_close(b.obs[1].cpu()[same], ref[same], rtol=H.RTOL, atol=2e-5)
    eng.close()


def test_height_scan_properties_flat_and_planar(native_lib):
    """Size-independent properties at the benchmark size: a flat field gives a constant scan for any yaw; a planar
    field is reproduced by both triangles of every cell."""
    cfg, spec = H.make_spec("go2_rough")
    n = 4096
    st = make_state(spec, n, seed=61)
    g = torch.Generator().manual_seed(62)
    st["root_pos_w"] = torch.stack([(torch.rand(n, generator=g) - 0.5) * 30.0, (torch.rand(n, generator=g) - 0.5) * 30.0,
                                    torch.rand(n, generator=g)], dim=1)
    starts = terrain_host.grid_pattern_ray_starts(RayCasterCfg()).cuda()
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n)
    b.load_logical(st)
    nx = ny = 400
    eng.height_scan_cast(b, terrain_host.HeightFieldBuffers(torch.full((nx, ny), 0.25, device="cuda"), -20.0, -20.0, 0.1, starts))
    assert (b.logical("ray_hits_z") == 0.25).all()
    xs = torch.arange(nx, dtype=torch.float64) * 0.1 - 20.0
    plane = (0.05 * xs[:, None] - 0.03 * xs[None, :] + 1.0).float().cuda()
    eng.height_scan_cast(b, terrain_host.HeightFieldBuffers(plane, -20.0, -20.0, 0.1, starts))
    torch.cuda.synchronize()
    q = st["root_quat_w"].double()
    yaw = torch.atan2(2 * (q[:, 0] * q[:, 3] + q[:, 1] * q[:, 2]), 1 - 2 * (q[:, 2] ** 2 + q[:, 3] ** 2))
    s = starts.cpu().double()
    p = st["root_pos_w"].double()
    wx = p[:, 0:1] + torch.cos(yaw)[:, None] * s[None, :, 0] - torch.sin(yaw)[:, None] * s[None, :, 1]
    wy = p[:, 1:2] + torch.sin(yaw)[:, None] * s[None, :, 0] + torch.cos(yaw)[:, None] * s[None, :, 1]
    torch.testing.assert_close(b.logical("ray_hits_z").cpu().double(), 0.05 * wx - 0.03 * wy + 1.0, rtol=0, atol=2e-5)
    eng.close()

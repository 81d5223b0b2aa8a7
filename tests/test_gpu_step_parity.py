"""GPU parity of the fused CUDA step against the CPU oracle (through the C-ABI, same seeded inputs).

Bar (BASELINE.json north_star): 1e-5 relative fp32 for floats (+1e-6 absolute floor, see helpers.py), bit-exact
for termination masks, done bits, episode lengths and reset indices.
"""

import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200 import _native as nat
from robot_lab_b200.synthetic import make_state

pytestmark = pytest.mark.gpu


def _engine(spec):
    from robot_lab_b200.engine import MdpStepEngine

    return MdpStepEngine(spec, "cuda:0")


def _run(key, n, layout="soa", cfg_launch=None, full_layout=False, seed=1234):
    cfg, spec = H.make_spec(key, full_layout=full_layout)
    st = make_state(spec, n, seed=seed)
    eng = _engine(spec)
    if cfg_launch:
        eng.set_launch_config(cfg_launch)
    b = eng.new_buffers(n, layout=layout)
    b.load_logical(st)
    eng.step(b)
    torch.cuda.synchronize()
    got, ref = H.gpu_step_outputs(b), H.oracle_step(spec, st)
    eng.close()
    return got, ref


# BASELINE.json configs: (0) A1 flat 64 envs, (1) Go2 flat 4096, (2) Go2 rough 4096, (3) G1 rough 4096 (J=29 per
# the reference, plus the J=37 label variant)
@pytest.mark.parametrize("key,n", [("a1_flat", 64), ("go2_flat", 4096), ("go2_rough", 4096), ("g1_rough", 4096),
                                   ("g1_rough_37", 4096), ("g1_flat", 1024), ("a1_rough", 1024)])
def test_full_step_matches_oracle(native_lib, key, n):
    got, ref = _run(key, n)
    H.compare_outputs(got, ref)
    assert len(ref["reset_ids"]) > 0, "synthetic state must exercise the reset path"


@pytest.mark.parametrize("n", [1, 3, 31, 33, 257, 4097])
def test_ragged_env_counts(native_lib, n):
    got, ref = _run("go2_rough", n)
    H.compare_outputs(got, ref)


@pytest.mark.parametrize("warps", [4, 8, 16])
@pytest.mark.parametrize("full_layout", [False, True])
def test_launch_configs_agree(native_lib, warps, full_layout):
    """Every warp count (different static schedules), baked kernels (compact layout) and the generic kernel."""
    got, ref = _run("go2_rough", 1000, cfg_launch=warps, full_layout=full_layout)
    H.compare_outputs(got, ref)


@pytest.mark.parametrize("key", ["go2_rough", "g1_rough"])
def test_isaaclab_shaped_aos_tensors(native_lib, key):
    """Every field handed over as [N, C] (PhysX-style) and every body in every body tensor."""
    got, ref = _run(key, 777, layout="aos", full_layout=True)
    H.compare_outputs(got, ref)


def test_skip_done_envs_then_refresh(native_lib):
    """Fused step with RL_PHASE_SKIP_DONE_ENVS + manager reset + COMMAND|OBS refresh of the reset ids ==
    the reference order: rewards -> reset -> command.compute -> observations."""
    cfg, spec = H.make_spec("go2_rough")
    n = 2048
    st = make_state(spec, n)
    eng = _engine(spec)
    b = eng.new_buffers(n)
    b.load_logical(st)
    rnd = H.rnd_inputs(st)
    eng.step(b, phases=nat.PHASE_ALL | nat.PHASE_SKIP_DONE_ENVS)
    torch.cuda.synchronize()
    got = H.gpu_step_outputs(b)
    ref = port.step(spec, st, rnd, skip_done_envs=True)
    H.compare_outputs(got, ref, keys=[k for k in ref if not k.startswith("obs_")])
    live = ~(ref["terminated"] | ref["truncated"])
    torch.testing.assert_close(got["obs_policy"][live], ref["obs_policy"][live], rtol=H.RTOL, atol=H.ATOL)
    # manager reset of the done envs
    eng.reset_envs(b, b.reset_ids, b.n_reset)
    torch.cuda.synchronize()
    st1 = dict(st)
    st1.update({k: ref[k] for k in ("command", "heading_target", "time_left", "is_heading_env", "is_standing_env",
                                    "metric_error_vel_xy", "metric_error_vel_yaw", "episode_length", "episode_sums")})
    st2, log = port.reset_envs(spec, st1, ref["reset_ids"], ref["done_bits"], rnd)
    for k in st2:
        g = b.logical(k).cpu().contiguous()
        if st2[k].dtype in (torch.bool, torch.int32):
            assert torch.equal(g.to(st2[k].dtype), st2[k]), k
        else:
            torch.testing.assert_close(g, st2[k], rtol=H.RTOL, atol=H.ATOL, msg=k)
    torch.testing.assert_close(b.log_episode_sum_mean[: spec.K].cpu(), log["episode_sum_mean"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(b.log_done_term_count.cpu(), log["done_term_count"], rtol=0, atol=0)
    torch.testing.assert_close(b.log_metric_mean.cpu(), log["metric_mean"], rtol=1e-4, atol=1e-6)
    # refresh: command.compute + observations for the reset ids only
    eng.step(b, phases=nat.PHASE_COMMAND | nat.PHASE_OBS, env_ids=b.reset_ids, n_env_ids=b.n_reset)
    torch.cuda.synchronize()
    st3 = dict(st1)
    st3.update(st2)
    ids = ref["reset_ids"].long()
    mask = torch.zeros(n, dtype=torch.bool)
    mask[ids] = True
    cmd3 = port.compute_command(spec, st3, rnd, active=mask)
    st4 = dict(st3)
    st4.update(cmd3)
    obs_p = port.compute_obs_group(spec, 0, st4, rnd)
    obs_c = port.compute_obs_group(spec, 1, st4, rnd)
    torch.testing.assert_close(b.obs[0].cpu()[ids], obs_p[ids], rtol=H.RTOL, atol=H.ATOL)
    torch.testing.assert_close(b.obs[1].cpu()[ids], obs_c[ids], rtol=H.RTOL, atol=H.ATOL)
    torch.testing.assert_close(b.obs[0].cpu()[live], ref["obs_policy"][live], rtol=H.RTOL, atol=H.ATOL)
    torch.testing.assert_close(b.logical("command").cpu().contiguous(), cmd3["command"], rtol=H.RTOL, atol=H.ATOL)
    eng.close()


@pytest.mark.parametrize("key,n", [("go2_rough", 2048), ("g1_rough", 333)])
def test_two_launch_step_in_reference_order(native_lib, key, n):
    """DONES|REWARDS|COMPACT, then RESET|COMMAND|OBS over all envs (reset applied to the flagged ones) ==
    terminations -> rewards -> _reset_idx -> command.compute -> observations, the order of ManagerBasedRLEnv.step()
    [IL], checked against the oracle run in that order."""
    cfg, spec = H.make_spec(key)
    st = make_state(spec, n)
    eng = _engine(spec)
    b = eng.new_buffers(n)
    b.load_logical(st)
    rnd = H.rnd_inputs(st)
    eng.step_pre_reset(b)
    torch.cuda.synchronize()
    ref = port.step(spec, st, rnd, skip_done_envs=True)   # dones / rewards / reset ids do not depend on the flag
    got = H.gpu_step_outputs(b)
    H.compare_outputs(got, ref, keys=["reward", "terminated", "truncated", "done_bits", "episode_length", "episode_sums",
                                      "step_reward", "reset_ids"])
    eng.step_post_reset(b)
    torch.cuda.synchronize()
    # oracle: reset the done envs, then command.compute and observations for everyone
    st1 = dict(st)
    st1.update({k: ref[k] for k in ("episode_length", "episode_sums")})
    st2, log = port.reset_envs(spec, st1, ref["reset_ids"], ref["done_bits"], rnd)
    st3 = dict(st1)
    st3.update(st2)
    cmd = port.compute_command(spec, st3, rnd, active=torch.ones(n, dtype=torch.bool))
    st4 = dict(st3)
    st4.update(cmd)
    obs_p = port.compute_obs_group(spec, 0, st4, rnd)
    obs_c = port.compute_obs_group(spec, 1, st4, rnd)
    torch.testing.assert_close(b.obs[0].cpu(), obs_p, rtol=H.RTOL, atol=H.ATOL)
    torch.testing.assert_close(b.obs[1].cpu(), obs_c, rtol=H.RTOL, atol=H.ATOL)
    for k in ("command", "heading_target", "time_left", "is_heading_env", "is_standing_env", "metric_error_vel_xy",
              "metric_error_vel_yaw", "episode_length", "episode_sums", "action", "prev_action"):
        want = st4[k]
        g = b.logical(k).cpu().contiguous()
        if want.dtype in (torch.bool, torch.int32):
            assert torch.equal(g.to(want.dtype), want), k
        else:
            torch.testing.assert_close(g, want, rtol=H.RTOL, atol=2e-6, msg=k)
    torch.testing.assert_close(b.log_episode_sum_mean[: spec.K].cpu(), log["episode_sum_mean"], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(b.log_done_term_count.cpu(), log["done_term_count"], rtol=0, atol=0)
    torch.testing.assert_close(b.log_metric_mean.cpu(), log["metric_mean"], rtol=1e-4, atol=1e-6)
    eng.close()


def test_process_action_matches_oracle(native_lib):
    for key in ("go2_rough", "g1_rough"):
        cfg, spec = H.make_spec(key)
        n = 513
        st = make_state(spec, n)
        eng = _engine(spec)
        b = eng.new_buffers(n)
        b.load_logical(st)
        eng.process_action(b)
        torch.cuda.synchronize()
        action, prev, processed = port.process_action(spec, st, st["new_action"])
        assert torch.equal(b.logical("action").cpu(), action)
        assert torch.equal(b.logical("prev_action").cpu(), prev)
        tgt = b.logical("joint_target").cpu()[:, spec.action.joint_ids]
        torch.testing.assert_close(tgt, processed, rtol=H.RTOL, atol=H.ATOL)
        eng.close()


def test_size_independent_properties_at_full_size(native_lib):
    """Properties that need no oracle, at BASELINE's full size: reward == sum_k step_reward_k * dt,
    episode sums advance by exactly the per-term values, reset ids == nonzero(done) ascending, masks disjoint
    from nothing-changed outputs, and a permutation of the envs permutes every output."""
    cfg, spec = H.make_spec("go2_rough")
    n = 4096
    st = make_state(spec, n)
    eng = _engine(spec)
    b = eng.new_buffers(n)
    b.load_logical(st)
    eng.step(b)
    torch.cuda.synchronize()
    got = H.gpu_step_outputs(b)
    dt = torch.tensor(spec.step_dt, dtype=torch.float32)
    acc = torch.zeros(n)
    for k in range(spec.K):  # same accumulation order as the manager
        acc += got["step_reward"][:, k] * dt
    torch.testing.assert_close(got["reward"], acc, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got["episode_sums"] - st["episode_sums"], got["step_reward"] * dt, rtol=1e-4, atol=2e-6)
    done = got["terminated"] | got["truncated"]
    assert torch.equal(got["reset_ids"].long(), done.nonzero().flatten())
    assert torch.equal(got["episode_length"], st["episode_length"] + 1)
    # permutation equivariance
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(7))
    st_p = {k: (v[:, perm] if k == "cmd_uniforms" else v[perm]) for k, v in st.items()}
    b2 = eng.new_buffers(n)
    b2.load_logical(st_p)
    eng.step(b2)
    torch.cuda.synchronize()
    got_p = H.gpu_step_outputs(b2)
    for k in ("reward", "obs_policy", "obs_critic", "terminated", "truncated", "command", "episode_sums"):
        assert torch.equal(got_p[k], got[k][perm]), k
    eng.close()


def test_production_philox_stream_matches_oracle(native_lib):
    """No random inputs given -> in-kernel Philox4x32-10; the oracle reproduces the same counters."""
    cfg, spec = H.make_spec("go2_rough")
    n = 1500
    st = make_state(spec, n)
    eng = _engine(spec)
    b = eng.new_buffers(n)
    b.load_logical(st)
    eng.step(b, seed=0x1234_5678_9ABC, step=41, env_id_offset=3 * n, use_random_inputs=False)
    torch.cuda.synchronize()
    got = H.gpu_step_outputs(b)
    rnd = {"seed": 0x1234_5678_9ABC, "step": 41, "env_id_offset": 3 * n}
    ref = port.step(spec, st, rnd)
    H.compare_outputs(got, ref)
    eng.close()


def test_empty_inputs_and_bad_arguments(native_lib):
    """num_envs == 0 is a no-op for every entry point; malformed calls come back as RL_EINVAL with a message, never as
    a CUDA error or a crash (the reference raises Python exceptions at the same places)."""
    import ctypes as C

    cfg, spec = H.make_spec("go2_rough")
    eng = _engine(spec)
    b = eng.new_buffers(64)
    b.load_logical(make_state(spec, 64))
    lib = eng.lib
    st, mdp, out, rnd = b.state_view(), b.mdp_state(), b.step_out(), b.random()
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.rl_step(eng._ctx, 0, C.byref(st), C.byref(mdp), C.byref(out), C.byref(rnd), nat.PHASE_ALL, None, None, stream) == 0
    na = b.field("new_action")
    assert lib.rl_process_action(eng._ctx, 0, C.byref(na), C.byref(mdp), None, None, None, stream) == 0
    # COMPACT without DONES, RESET together with REWARDS, RESET without the masks it needs, a missing observation input
    for phases in (nat.PHASE_COMPACT, nat.PHASE_RESET | nat.PHASE_REWARDS):
        with pytest.raises(nat.NativeError):
            eng.step(b, phases=phases)
    bad = b.step_out(fresh=True)
    bad.terminated = None
    with pytest.raises(nat.NativeError):
        nat.check(lib.rl_step(eng._ctx, 64, C.byref(st), C.byref(mdp), C.byref(bad), C.byref(rnd),
                              nat.PHASE_RESET | nat.PHASE_COMMAND | nat.PHASE_OBS, None, None, stream))
    st2 = b.state_view(fresh=True)
    st2.ray_hits_z = nat.RlField(None, 0, 0)
    with pytest.raises(nat.NativeError):
        nat.check(lib.rl_step(eng._ctx, 64, C.byref(st2), C.byref(mdp), C.byref(out), C.byref(rnd), nat.PHASE_OBS, None, None, stream))
    torch.cuda.synchronize()
    eng.step(b)   # the context is still usable afterwards
    torch.cuda.synchronize()
    eng.close()


def test_two_action_terms_position_and_velocity_targets(native_lib):
    """JointPositionAction + JointVelocityAction in one action vector (the wheeled robots' ActionsCfg,
    V/config/wheeled/unitree_go2w/rough_env_cfg.py:22-32): position columns land in joint_target, velocity columns in
    joint_vel_target, the stored action / last_action observation carry all of them."""
    from test_spec_compile import _two_term_action_cfg

    from robot_lab_b200.spec import compact_layout, compile_step_spec

    cfg, legs, wheels = _two_term_action_cfg()
    spec = compile_step_spec(cfg, compact_layout(cfg))
    n = 700
    st = make_state(spec, n)
    eng = _engine(spec)
    b = eng.new_buffers(n)
    b.load_logical(st)
    b.t["joint_target"].fill_(-7.0)
    b.t["joint_vel_target"].fill_(-9.0)
    eng.process_action(b)
    torch.cuda.synchronize()
    action, prev, processed = port.process_action(spec, st, st["new_action"])
    assert torch.equal(b.logical("action").cpu(), action) and torch.equal(b.logical("prev_action").cpu(), prev)
    pos_ids, vel_ids = spec.action.joint_ids[:8], spec.action.joint_ids[8:]
    jt, jv = b.logical("joint_target").cpu(), b.logical("joint_vel_target").cpu()
    torch.testing.assert_close(jt[:, pos_ids], processed[:, :8], rtol=H.RTOL, atol=H.ATOL)
    torch.testing.assert_close(jv[:, vel_ids], processed[:, 8:], rtol=H.RTOL, atol=H.ATOL)
    assert (jt[:, vel_ids] == -7.0).all() and (jv[:, pos_ids] == -9.0).all()     # columns not driven are untouched
    # the full step (generic kernel: this spec is not one of the baked ones) against the oracle
    st2 = dict(st)
    st2["action"], st2["prev_action"] = action, prev
    eng.step(b)
    torch.cuda.synchronize()
    H.compare_outputs(H.gpu_step_outputs(b), H.oracle_step(spec, st2))
    eng.close()

"""The env / manager / wrapper surface the reference's scripts consume, on the GPU."""

import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200.synthetic import MDP_KEYS, make_state

pytestmark = pytest.mark.gpu


def test_zero_agent_loop(native_lib):
    """scripts/tools/zero_agent.py:56-70: make -> reset -> step(zeros) repeatedly."""
    from robot_lab_b200.envs import make

    env = make(H.TASKS["go2_rough"], num_envs=512, device="cuda:0")
    obs, extras = env.reset()
    assert obs["policy"].shape == (512, 45) and obs["critic"].shape == (512, 235)
    assert env.action_space.shape == (512, 12) and env.unwrapped is env
    ep0 = env.episode_length_buf.clone()
    assert ep0.dtype == torch.int64 and int(ep0.max()) == 0
    for _ in range(20):
        obs, rew, terminated, truncated, extras = env.step(torch.zeros(512, 12, device="cuda:0"))
    torch.cuda.synchronize()
    assert rew.shape == (512,) and terminated.dtype == torch.bool and truncated.dtype == torch.bool
    assert torch.isfinite(obs["policy"]).all() and torch.isfinite(obs["critic"]).all() and torch.isfinite(rew).all()
    assert (obs["critic"][:, 48:].abs() <= 1.0).all()  # height scan clipped to +-1
    assert "Episode_Reward/track_lin_vel_xy_exp" in extras["log"] and "Episode_Termination/time_out" in extras["log"]
    done = terminated | truncated
    ep = env.episode_length_buf
    assert (ep[done] == 0).all() and (ep[~done] > 0).all()
    # last_action observation = the action just applied (zeros), for every env
    assert (obs["policy"][:, 33:45] == 0).all()
    env.close()


def test_term_function_protocol(native_lib):
    """``func(env, **params) -> Tensor[N]`` still works for a single term (V/mdp/rewards.py:22 ff.)."""
    from robot_lab_b200 import mdp
    from robot_lab_b200.cfg import SceneEntityCfg
    from robot_lab_b200.envs import make

    env = make(H.TASKS["go2_rough"], num_envs=256, device="cuda:0")
    env.reset()
    env.step(torch.randn(256, 12, device="cuda:0"))
    v = mdp.track_lin_vel_xy_exp(env, std=0.5, command_name="base_velocity")
    w = mdp.feet_slide(env, sensor_cfg=SceneEntityCfg("contact_forces", body_names=[".*_foot"]),
                       asset_cfg=SceneEntityCfg("robot", body_names=[".*_foot"]))
    torch.cuda.synchronize()
    assert v.shape == (256,) and w.shape == (256,) and (v >= 0).all() and (v <= 1).all() and (w >= 0).all()
    assert env.scene["robot"].data.joint_pos.shape == (256, 12)
    assert env.command_manager.get_command("base_velocity").shape == (256, 3)
    assert env.action_manager.action.shape == (256, 12) and env.reward_manager._episode_sums["upward"].shape == (256,)
    env.close()


def test_env_rollout_matches_oracle_replay(native_lib):
    """Stateful semantics over several steps: episode sums, command timers, resets, refreshed observations."""
    from robot_lab_b200.envs import ManagerBasedRLEnv, ReplayStateProvider
    from robot_lab_b200.tasks import make_env_cfg

    n, T, seed = 768, 5, 77
    cfg = make_env_cfg(H.TASKS["go2_rough"], num_envs=n)
    cfg.seed, cfg.sim.device = seed, "cuda:0"
    _, spec = H.make_spec("go2_rough")
    phys = [make_state(spec, n, seed=500 + t) for t in range(T)]
    actions = [torch.randn(n, spec.A, generator=torch.Generator().manual_seed(t)) for t in range(T)]
    env = ManagerBasedRLEnv(cfg, state_provider=ReplayStateProvider(phys))
    # start from a non-trivial manager state instead of reset(): load it straight into the buffers
    mdp0 = {k: phys[0][k] for k in MDP_KEYS}
    env.buffers.load_logical(mdp0)
    env.buffers.cmd_uniforms, env.buffers.obs_uniforms = None, [None, None]
    got = []
    for t in range(T):
        obs, rew, term, trunc, _ = env.step(actions[t].cuda())
        torch.cuda.synchronize()
        got.append({"obs_policy": obs["policy"].cpu().clone(), "obs_critic": obs["critic"].cpu().clone(),
                    "reward": rew.cpu().clone(), "terminated": term.cpu(), "truncated": trunc.cpu(),
                    "mdp": {k: env.buffers.logical(k).cpu().clone().contiguous() for k in MDP_KEYS}})
    ref = H.oracle_env_rollout(spec, mdp0, phys, actions, seed, n)
    for t in range(T):
        assert torch.equal(got[t]["terminated"], ref[t]["terminated"]) and torch.equal(got[t]["truncated"], ref[t]["truncated"]), t
        for k in ("obs_policy", "obs_critic", "reward"):
            torch.testing.assert_close(got[t][k], ref[t][k], rtol=H.RTOL, atol=H.ATOL, msg=f"step {t} {k}")
        for k in MDP_KEYS:
            g, r = got[t]["mdp"][k], ref[t]["mdp"][k]
            if r.dtype in (torch.bool, torch.int32):
                assert torch.equal(g.to(r.dtype), r), (t, k)
            else:
                torch.testing.assert_close(g, r, rtol=H.RTOL, atol=2e-6, msg=f"step {t} {k}")
    env.close()


def test_rsl_rl_wrapper_surface(native_lib):
    """What OnPolicyRunner touches (scripts/reinforcement_learning/rsl_rl/train.py:202-224)."""
    from robot_lab_b200.envs import RslRlVecEnvWrapper, make

    env = RslRlVecEnvWrapper(make(H.TASKS["go2_flat"], num_envs=128, device="cuda:0"), clip_actions=1.0)
    assert (env.num_envs, env.num_actions, env.max_episode_length) == (128, 12, 1000)
    obs = env.get_observations()
    assert obs["policy"].shape == (128, 45) and obs["critic"].shape == (128, 48)
    # init_at_random_ep_len=True: rsl_rl overwrites the episode length buffer
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
    ep = env.episode_length_buf.clone()
    obs, rew, dones, extras = env.step(5.0 * torch.ones(128, 12, device="cuda:0"))
    torch.cuda.synchronize()
    assert dones.dtype == torch.long and "time_outs" in extras and "log" in extras
    live = dones == 0
    act = env.unwrapped.action_manager.action
    assert (act[live] == 1.0).all()            # clipped to clip_actions
    assert (act[~live] == 0.0).all()           # ActionManager.reset for the envs that were reset
    assert torch.equal(env.episode_length_buf[live], ep[live] + 1)
    env.close()


def test_env_step_on_a_terrain_with_pits(native_lib):
    """A terrain with a "pits" sub-terrain switches env.step() to command -> rl_command_pit_restrict -> observations
    (V/mdp/commands.py:61-85): the restricted command is what the policy observes."""
    from robot_lab_b200 import terrain as terrain_host
    from robot_lab_b200.cfg import TerrainCfg
    from robot_lab_b200.envs import ManagerBasedRLEnv, ReplayStateProvider
    from robot_lab_b200.tasks import make_env_cfg

    n = 1024
    cfg = make_env_cfg(H.TASKS["go2_rough"], n)
    cfg.sim.device = "cuda:0"
    cfg.scene.terrain = TerrainCfg(sub_terrains=("pyramid_stairs", "pits", "boxes", "random_rough", "hf_pyramid_slope"),
                                   proportions=(0.2, 0.15, 0.25, 0.3, 0.1))
    _, spec = H.make_spec("go2_rough")
    g = torch.Generator().manual_seed(8)
    states = []
    for i in range(4):
        st = make_state(spec, n, seed=70 + i)
        st["root_pos_w"] = torch.stack([(torch.rand(n, generator=g) - 0.5) * 70.0, (torch.rand(n, generator=g) - 0.5) * 110.0,
                                        torch.rand(n, generator=g)], dim=1)   # inside the terrain: no out-of-bounds resets
        states.append(st)
    env = ManagerBasedRLEnv(cfg, state_provider=ReplayStateProvider(states))
    assert env.pit_grid is not None and (env.pit_grid.col_start, env.pit_grid.col_end) == (4, 7)
    env.reset()
    origins = terrain_host.grid_origins(cfg.scene.terrain)
    prev_on = None
    for i in range(1, 4):
        obs, rew, terminated, truncated, _ = env.step(torch.zeros(n, 12, device="cuda:0"))
        torch.cuda.synchronize()
        on = port.is_robot_on_terrain(states[i]["root_pos_w"], origins, (4, 7))
        assert torch.equal(env.was_on_pit.cpu().bool(), on) and 50 < int(on.sum()) < n // 2
        cmd = env.command_manager.get_command("base_velocity").cpu()
        assert ((cmd[on, 0] >= 0.3) & (cmd[on, 0] <= 0.6)).all() and (cmd[on, 1:] == 0).all()
        assert torch.equal(obs["policy"][:, 6:9].cpu(), cmd)      # generated_commands columns of the policy row
        if prev_on is not None:
            left = prev_on & ~on
            assert int(left.sum()) > 10   # ... whose commands were resampled: not all of them forward-only any more
            assert (cmd[left, 1] != 0).any()
        prev_on = on
    env.close()


def test_sensor_chain_state_provider(native_lib):
    """env.step() with every neighbour of the path computed on the device: actuator model, contact sensor, height
    scanner, reset events."""
    from robot_lab_b200 import terrain as terrain_host
    from robot_lab_b200.cfg import RayCasterCfg
    from robot_lab_b200.envs import ManagerBasedRLEnv, SensorChainStateProvider
    from robot_lab_b200.tasks import make_env_cfg

    n = 512
    cfg = make_env_cfg(H.TASKS["go2_rough"], n)
    cfg.sim.device = "cuda:0"
    _, spec = H.make_spec("go2_rough")
    nx = ny = 1400
    g = torch.Generator().manual_seed(9)
    heights = (torch.rand(nx, ny, generator=g) * 0.1).cuda()
    hf = terrain_host.HeightFieldBuffers(heights, -70.0, -70.0, 0.1, terrain_host.grid_pattern_ray_starts(RayCasterCfg()).cuda())
    prov = SensorChainStateProvider(spec, n, "cuda:0", height_field=hf, num_sets=2)
    env = ManagerBasedRLEnv(cfg, state_provider=prov)
    env.reset()
    act = torch.randn(n, 12, device="cuda:0")
    obs, rew, terminated, truncated, _ = env.step(act)
    torch.cuda.synchronize()
    b = env.buffers
    tab = spec.layout.asset.actuator_table()
    _, want = port.actuator_step(tab, b.logical("joint_target").cpu().contiguous(), b.logical("joint_pos").cpu().contiguous(),
                                 b.logical("joint_vel").cpu().contiguous())
    done = (terminated | truncated).cpu()
    torch.testing.assert_close(b.logical("applied_torque").cpu().contiguous()[~done], want[~done], rtol=H.RTOL, atol=H.ATOL)
    hits = b.logical("ray_hits_z").cpu()
    # (the scan is taken before the reset events move the done envs, like IsaacLab's lazily updated sensors)
    inside = (b.logical("root_pos_w").cpu()[:, :2].abs() < 65.0).all(dim=1) & ~done
    assert torch.isfinite(hits[inside]).all() and (hits[inside] >= 0).all() and (hits[inside] <= 0.1).all()
    t = b.logical("current_air_time").cpu() + b.logical("current_contact_time").cpu()
    assert (t > 0).all()      # four sub-steps of 5 ms were accumulated on one of the two timers of every foot
    assert torch.isfinite(obs["critic"]).all() and torch.isfinite(rew).all()
    env.close()

"""SURVEY.md 8(b).2 on the device: the derived articulation views (rl_derived_views) against the oracle and - bit for
bit - against what the step kernels compute from the same state; term functions written against the env surface the
way the reference's are (V/mdp/rewards.py:22-35, :557-587) against rl_term_eval; the sensors / terrain objects.
(The UNMODIFIED reference functions run against the same scene classes in tests/test_env_surface_reference_terms.py,
in the build container where /root/reference exists.)"""

import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200.synthetic import make_state

pytestmark = pytest.mark.gpu


def _env(key="go2_rough", n=1024, body_tensors="full"):
    from robot_lab_b200 import envs
    from robot_lab_b200.tasks import make_env_cfg

    cfg = make_env_cfg(H.TASKS[key], num_envs=n)
    cfg.sim.device = "cuda:0"
    env = envs.ManagerBasedRLEnv(cfg, body_tensors=body_tensors)
    st = make_state(env.spec, n, seed=31)
    env.buffers.load_logical(st)
    env.invalidate_derived()
    return env, st


def test_derived_views_match_oracle_and_the_step_kernels(native_lib):
    env, st = _env(body_tensors="compact")
    d = port.Derived(st, env.spec)
    data = env.scene["robot"].data
    for name, want in (("projected_gravity_b", d.projected_gravity_b), ("root_lin_vel_b", d.root_lin_vel_b),
                       ("root_ang_vel_b", d.root_ang_vel_b), ("root_com_lin_vel_b", d.root_lin_vel_b), ("heading_w", d.heading_w())):
        torch.testing.assert_close(getattr(data, name).cpu(), want, rtol=1e-5, atol=2e-6, msg=name)
    # the critic row (no noise, clip +-100, scale 1) starts with base_lin_vel, base_ang_vel, projected_gravity: the very
    # same arithmetic inside the step kernel -> bit-identical
    env.observation_manager.compute()
    torch.cuda.synchronize()
    critic = env.buffers.obs[1]
    names = [t.type_name for t in env.spec.obs[1].terms]
    col = 0
    for t in env.spec.obs[1].terms:
        view = {"base_lin_vel": data.root_lin_vel_b, "base_ang_vel": data.root_ang_vel_b, "projected_gravity": data.projected_gravity_b}.get(t.type_name)
        if view is not None and t.scale in (None, 1.0) and (t.clip is None or abs(t.clip[0]) >= 100):
            assert torch.equal(critic[:, col:col + 3], view), t.type_name
        col += t.dim
    assert "projected_gravity" in names
    # constants and aliases
    torch.testing.assert_close(data.default_joint_pos[0].cpu(), torch.tensor(env.spec.default_joint_pos), rtol=0, atol=0)
    assert tuple(data.soft_joint_pos_limits.shape) == (env.num_envs, env.spec.J, 2)
    assert data.root_link_quat_w.data_ptr() == data.root_quat_w.data_ptr()
    # the views follow the state: new physical state + invalidate -> new values
    st2 = make_state(env.spec, env.num_envs, seed=77)
    env.buffers.load_logical(st2)
    env.invalidate_derived()
    torch.testing.assert_close(data.projected_gravity_b.cpu(), port.Derived(st2, env.spec).projected_gravity_b, rtol=1e-5, atol=2e-6)
    env.close()


def track_lin_vel_xy_exp_on_surface(env, std, command_name, asset_cfg):
    """The op sequence of V/mdp/rewards.py:22-35, against the env surface."""
    asset = env.scene[asset_cfg.name]
    err = torch.sum(torch.square(env.command_manager.get_command(command_name)[:, :2] - asset.data.root_lin_vel_b[:, :2]), dim=1)
    reward = torch.exp(-err / std**2)
    return reward * (torch.clamp(-asset.data.projected_gravity_b[:, 2], 0, 0.7) / 0.7)


def feet_slide_on_surface(env, sensor_cfg, asset_cfg):
    """The op sequence of V/mdp/rewards.py:557-587, against the env surface."""
    sensor = env.scene.sensors[sensor_cfg.name]
    contacts = sensor.data.net_forces_w_history[:, :, sensor_cfg.body_ids, :].norm(dim=-1).max(dim=1)[0] > 1.0
    asset = env.scene[asset_cfg.name]
    rel = asset.data.body_lin_vel_w[:, asset_cfg.body_ids, :] - asset.data.root_lin_vel_w.unsqueeze(1)
    q = asset.data.root_quat_w
    in_body = torch.stack([port.quat_apply_inverse(q.cpu(), rel[:, i].cpu()) for i in range(rel.shape[1])], dim=1).to(q.device)
    lat = torch.sqrt(torch.sum(torch.square(in_body[:, :, :2]), dim=2))
    reward = torch.sum(lat * contacts, dim=1)
    return reward * (torch.clamp(-env.scene["robot"].data.projected_gravity_b[:, 2], 0, 0.7) / 0.7)


def test_surface_terms_match_rl_term_eval(native_lib):
    from robot_lab_b200 import mdp
    from robot_lab_b200.cfg import SceneEntityCfg

    env, st = _env()
    asset_cfg = SceneEntityCfg("robot")
    asset_cfg.resolve(env.scene)
    got = track_lin_vel_xy_exp_on_surface(env, std=0.5, command_name="base_velocity", asset_cfg=asset_cfg)
    want = env.reward_manager.evaluate_term(mdp.track_lin_vel_xy_exp, {"std": 0.5, "command_name": "base_velocity"})
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    sensor_cfg = SceneEntityCfg("contact_forces", body_names=".*_foot")
    feet_cfg = SceneEntityCfg("robot", body_names=".*_foot")
    sensor_cfg.resolve(env.scene)
    feet_cfg.resolve(env.scene)
    got = feet_slide_on_surface(env, sensor_cfg, feet_cfg)
    want = env.reward_manager.evaluate_term(mdp.feet_slide, {"sensor_cfg": SceneEntityCfg("contact_forces", body_names=".*_foot"),
                                                             "asset_cfg": SceneEntityCfg("robot", body_names=".*_foot")})
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    env.close()


def test_sensors_terrain_and_origins(native_lib):
    env, st = _env(n=512)
    hs = env.scene.sensors["height_scanner"]
    assert tuple(hs.data.pos_w.shape) == (512, 3) and tuple(hs.data.ray_hits_w.shape) == (512, env.spec.R, 3)
    want = st["ray_sensor_pos_z"].reshape(-1, 1) - st["ray_hits_z"] - 0.5
    torch.testing.assert_close((hs.data.pos_w[:, 2:3] - hs.data.ray_hits_w[..., 2] - 0.5).cpu(), want, rtol=0, atol=0)
    ter = env.scene.terrain
    assert ter.cfg.terrain_type == "generator" and set(ter.cfg.terrain_generator.sub_terrains) >= {"boxes", "random_rough"}
    assert tuple(ter.terrain_origins.shape) == (10, 20, 3) and tuple(env.scene.env_origins.shape) == (512, 3)
    assert ter.terrain_types.shape == (512,) and int(ter.terrain_types.max()) <= 19
    cs = env.scene.sensors["contact_forces"]
    assert torch.equal(cs.data.net_forces_w, cs.data.net_forces_w_history[:, 0])
    assert cs.compute_first_contact(env.step_dt).shape == cs.data.current_contact_time.shape
    env.close()


def test_extras_log_is_lazy_and_survives_later_steps(native_lib):
    """rsl_rl keeps infos["log"] of every step and reads them at the end of the iteration."""
    from robot_lab_b200 import envs
    from robot_lab_b200.tasks import make_env_cfg

    cfg = make_env_cfg(H.TASKS["go2_rough"], num_envs=256)
    cfg.sim.device = "cuda:0"
    env = envs.RslRlVecEnvWrapper(envs.ManagerBasedRLEnv(cfg))
    logs, sums = [], []
    for t in range(3):
        obs, rew, dones, extras = env.step(torch.randn(256, env.num_actions, device="cuda:0"))
        logs.append(extras["log"])
        sums.append(env.unwrapped.buffers.log_all.clone())
    kk = max(env.unwrapped.spec.K, 1)
    for log, snap in zip(logs, sums):   # read late: each log still shows ITS step
        name = env.unwrapped.reward_manager.active_terms[0]
        assert torch.equal(log[f"Episode_Reward/{name}"], snap[0] / env.unwrapped.max_episode_length_s)
        assert len(log) == len(env.unwrapped.reward_manager.active_terms) + len(env.unwrapped.termination_manager.active_terms) + 2
        assert torch.equal(log["Metrics/base_velocity/error_vel_xy"], snap[kk + 8])
    env.close()

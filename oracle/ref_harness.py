"""Run the reference's OWN term functions (unmodified files under /root/reference) on a synthetic state.

TEST INFRASTRUCTURE, build-container only (needs /root/reference). A duck-typed fake env exposes exactly the
attributes the terms touch: ``env.scene["robot"].data.*``, ``env.scene.sensors[...]``,
``env.command_manager.get_command``, ``env.action_manager.action/prev_action``, ``env.num_envs/device/step_dt``.
The [IL]-owned derived quantities behind those attributes (projected gravity, body-frame velocities,
``compute_first_contact``) come from the restatements in ``oracle/mdp_port.py``.
"""

from __future__ import annotations

import sys
import types
from typing import Any

import torch

from robot_lab_b200.cfg import RewardTermCfg, SceneEntityCfg, resolve_matching_names
from robot_lab_b200.spec import (
    ASSET_SPACE_FIELDS,
    HIST_SPACE_FIELDS,
    TIME_SPACE_FIELDS,
    RewardTermSpec,
    StepSpec,
)

from . import isaaclab_shim as shim
from . import mdp_port as port


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Scene(dict):
    pass


class FakeEnv:
    def __init__(self, spec: StepSpec, st: dict):
        d = port.Derived(st, spec)
        n = d.N
        asset = _NS()
        asset.data = _NS(
            projected_gravity_b=d.projected_gravity_b,
            root_quat_w=st["root_quat_w"], root_link_quat_w=st["root_quat_w"],
            root_pos_w=st["root_pos_w"], root_link_pos_w=st["root_pos_w"],
            root_lin_vel_b=d.root_lin_vel_b, root_com_lin_vel_b=d.root_lin_vel_b,
            root_lin_vel_w=st["root_lin_vel_w"], root_ang_vel_b=d.root_ang_vel_b, root_ang_vel_w=st["root_ang_vel_w"],
            joint_pos=st["joint_pos"], joint_vel=st["joint_vel"], joint_acc=st["joint_acc"],
            default_joint_pos=d.default_joint_pos, default_joint_vel=d.default_joint_vel,
            applied_torque=st["applied_torque"],
            body_pos_w=st["body_pos_w"], body_link_pos_w=st["body_pos_w"], body_lin_vel_w=st["body_lin_vel_w"],
            heading_w=d.heading_w(),
        )
        jn = list(spec.layout.asset.joint_names)
        asset.find_joints = lambda keys, preserve_order=False: resolve_matching_names(keys, jn, preserve_order)
        asset.device = "cpu"
        sensor = _NS()
        sensor.data = _NS(
            net_forces_w_history=st["net_forces_w_history"], net_forces_w=st["net_forces_w_history"][:, 0],
            current_air_time=st["current_air_time"], last_air_time=st["last_air_time"],
            current_contact_time=st["current_contact_time"], last_contact_time=st["last_contact_time"],
        )
        sensor.compute_first_contact = lambda dt, abs_tol=1.0e-8: (
            (st["current_contact_time"] > 0.0) * (st["current_contact_time"] < (dt + abs_tol)))
        sensor.compute_first_air = lambda dt, abs_tol=1.0e-8: (
            (st["current_air_time"] > 0.0) * (st["current_air_time"] < (dt + abs_tol)))
        tn = list(spec.layout.time_body_names)
        sensor.find_bodies = lambda keys, preserve_order=False: resolve_matching_names(keys, tn, preserve_order)
        self.scene = _Scene(robot=asset, contact_forces=sensor)
        self.scene.sensors = {"contact_forces": sensor}
        self.scene.terrain = None
        self.scene.num_envs = n
        self.num_envs = n
        self.device = "cpu"
        self.step_dt = spec.step_dt
        self.episode_length_buf = st["episode_length"].long()
        self.command_manager = _NS(get_command=lambda name: st["command"])
        self.action_manager = _NS(action=st["action"], prev_action=st["prev_action"])


def _shim_entity(name: str, joint_ids=slice(None), body_ids=slice(None)):
    return shim.SceneEntityCfg(name, joint_ids=joint_ids, body_ids=body_ids)


def reference_params(cfg: RewardTermCfg, t: RewardTermSpec) -> dict[str, Any]:
    """cfg.params with every SceneEntityCfg replaced by one holding the ids the spec compiler resolved."""
    params = dict(cfg.func.rl_defaults)
    params.update(cfg.params)
    ty = t.type_name
    out = {}
    for k, v in params.items():
        if isinstance(v, SceneEntityCfg):
            if k == "sensor_cfg":
                if ty in TIME_SPACE_FIELDS:
                    ids = getattr(t, TIME_SPACE_FIELDS[ty][0])
                elif ty in HIST_SPACE_FIELDS:
                    ids = getattr(t, HIST_SPACE_FIELDS[ty][0])
                else:
                    ids = slice(None)
                out[k] = _shim_entity(v.name, body_ids=ids)
            else:  # asset_cfg
                body = getattr(t, ASSET_SPACE_FIELDS[ty][0]) if ty in ASSET_SPACE_FIELDS else slice(None)
                joints = t.joint_ids if t.joint_ids else (t.idx_b if ty == "wheel_vel_penalty" else slice(None))
                out[k] = _shim_entity(v.name, joint_ids=joints, body_ids=body)
        else:
            out[k] = v
    return out


def reference_reward_terms(env_cfg, spec: StepSpec, st: dict) -> dict[str, torch.Tensor | None]:
    """name -> raw term value from the reference function, or None when the term is IsaacLab-owned."""
    ref = shim.load_reference_module("rewards.py")
    env = FakeEnv(spec, st)
    by_name = {t.name: t for t in spec.rewards}
    out: dict[str, torch.Tensor | None] = {}
    for name, cfg in env_cfg.rewards.active(RewardTermCfg):
        t = by_name[name]
        fname = cfg.func.__name__
        if not hasattr(ref, fname):
            out[name] = None
            continue
        fn = getattr(ref, fname)
        params = reference_params(cfg, t)
        if isinstance(fn, type):  # class term (GaitReward): ManagerTermBase(cfg, env).__call__
            term = fn(shim.RewardTermCfg(func=fn, weight=cfg.weight, params=params), env)
            out[name] = term(env, **params)
        else:
            out[name] = fn(env, **params)
    return out


def reference_observation_terms(spec: StepSpec, st: dict) -> dict[str, torch.Tensor]:
    """The two robot_lab-owned observation functions (V/mdp/observations.py:17-35)."""
    ref = shim.load_reference_module("observations.py")
    env = FakeEnv(spec, st)
    n_j = spec.J
    wheel = [j for j in range(n_j) if j % 3 == 2]
    return {
        "phase": ref.phase(env, cycle_time=0.8),
        "joint_pos_rel_without_wheel": ref.joint_pos_rel_without_wheel(
            env, asset_cfg=_shim_entity("robot"), wheel_asset_cfg=_shim_entity("robot", joint_ids=wheel)),
        "_wheel_ids": torch.tensor(wheel),
    }


# ------------------------------------------------------------------------------------------------
# command term: the reference's UniformThresholdVelocityCommand over a restated [IL] base class
# ------------------------------------------------------------------------------------------------
def _install_command_context(uniform_table: torch.Tensor):
    """Fake package context so V/mdp/commands.py (relative + self imports, :14-16) loads unmodified."""
    shim.install()
    managers = sys.modules["isaaclab.managers"]

    class CommandTermCfg:
        pass

    class CommandTerm:  # CommandTerm [IL]: compute / _resample / reset skeleton
        def __init__(self, cfg, env):
            self.cfg, self._env = cfg, env
            self.num_envs, self.device = env.num_envs, env.device
            self.metrics = {}
            self.time_left = torch.zeros(self.num_envs)
            self.command_counter = torch.zeros(self.num_envs, dtype=torch.long)

        def compute(self, dt: float):
            self._update_metrics()
            self.time_left -= dt
            ids = (self.time_left <= 0.0).nonzero().flatten()
            if len(ids) > 0:
                self._resample(ids)
            self._update_command()

        def _resample(self, env_ids):
            if len(env_ids) != 0:
                lo, hi = self.cfg.resampling_time_range
                self.time_left[env_ids] = self._u[0, env_ids] * (hi - lo) + lo
                self._resample_command(env_ids)
                self.command_counter[env_ids] += 1

    managers.CommandTerm, managers.CommandTermCfg = CommandTerm, CommandTermCfg
    sys.modules["isaaclab.utils"].configclass = lambda cls: cls

    class UniformVelocityCommandCfg(CommandTermCfg):
        pass

    class UniformVelocityCommand(CommandTerm):  # UniformVelocityCommand [IL], uniforms taken from a table
        def __init__(self, cfg, env):
            super().__init__(cfg, env)
            self.robot = env.scene[cfg.asset_name]
            n = self.num_envs
            self.vel_command_b = torch.zeros(n, 3)
            self.heading_target = torch.zeros(n)
            self.is_heading_env = torch.zeros(n, dtype=torch.bool)
            self.is_standing_env = torch.zeros_like(self.is_heading_env)
            self.metrics["error_vel_xy"] = torch.zeros(n)
            self.metrics["error_vel_yaw"] = torch.zeros(n)
            self._u = uniform_table

        @property
        def command(self):
            return self.vel_command_b

        def _update_metrics(self):
            max_command_step = self.cfg.resampling_time_range[1] / self._env.step_dt
            self.metrics["error_vel_xy"] += (
                torch.norm(self.vel_command_b[:, :2] - self.robot.data.root_lin_vel_b[:, :2], dim=-1) / max_command_step)
            self.metrics["error_vel_yaw"] += (
                torch.abs(self.vel_command_b[:, 2] - self.robot.data.root_ang_vel_b[:, 2]) / max_command_step)

        def _resample_command(self, env_ids):
            r, u = self.cfg.ranges, self._u
            self.vel_command_b[env_ids, 0] = u[1, env_ids] * (r.lin_vel_x[1] - r.lin_vel_x[0]) + r.lin_vel_x[0]
            self.vel_command_b[env_ids, 1] = u[2, env_ids] * (r.lin_vel_y[1] - r.lin_vel_y[0]) + r.lin_vel_y[0]
            self.vel_command_b[env_ids, 2] = u[3, env_ids] * (r.ang_vel_z[1] - r.ang_vel_z[0]) + r.ang_vel_z[0]
            if self.cfg.heading_command:
                self.heading_target[env_ids] = u[4, env_ids] * (r.heading[1] - r.heading[0]) + r.heading[0]
                self.is_heading_env[env_ids] = u[5, env_ids] <= self.cfg.rel_heading_envs
            self.is_standing_env[env_ids] = u[6, env_ids] <= self.cfg.rel_standing_envs

        def _update_command(self):
            if self.cfg.heading_command:
                ids = self.is_heading_env.nonzero(as_tuple=False).flatten()
                err = port.wrap_to_pi(self.heading_target[ids] - self.robot.data.heading_w[ids])
                self.vel_command_b[ids, 2] = torch.clip(
                    self.cfg.heading_control_stiffness * err, min=self.cfg.ranges.ang_vel_z[0], max=self.cfg.ranges.ang_vel_z[1])
            sid = self.is_standing_env.nonzero(as_tuple=False).flatten()
            self.vel_command_b[sid, :] = 0.0

    pkg = "robot_lab.tasks.manager_based.locomotion.velocity.mdp"
    parts = pkg.split(".")
    for i in range(1, len(parts) + 1):
        name = ".".join(parts[:i])
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    mdp_pkg = sys.modules[pkg]
    mdp_pkg.UniformVelocityCommand = UniformVelocityCommand
    mdp_pkg.UniformVelocityCommandCfg = UniformVelocityCommandCfg
    import importlib.util

    for fname in ("utils", "commands"):
        full = f"{pkg}.{fname}"
        sp = importlib.util.spec_from_file_location(full, shim.MDP_DIR / f"{fname}.py")
        module = importlib.util.module_from_spec(sp)
        module.__package__ = pkg
        sys.modules[full] = module
        sp.loader.exec_module(module)
        setattr(mdp_pkg, fname, module)
        if fname == "commands":
            mdp_pkg.UniformThresholdVelocityCommandCfg = module.UniformThresholdVelocityCommandCfg
    return sys.modules[f"{pkg}.commands"]


def fake_terrain(terrain_cfg, terrain_type: str, num_envs: int, terrain_origins: torch.Tensor | None = None,
                 terrain_types: torch.Tensor | None = None):
    """A terrain importer the way V/mdp/utils.py inspects it (:56-66, :86-107): cfg.terrain_type,
    cfg.terrain_generator.{sub_terrains{name: .proportion}, num_cols}, terrain_types, terrain_origins."""
    props = getattr(terrain_cfg, "proportions", None) or [1.0] * len(terrain_cfg.sub_terrains)
    gen = _NS(sub_terrains={k: _NS(proportion=float(p)) for k, p in zip(terrain_cfg.sub_terrains, props)},
              num_cols=terrain_cfg.num_cols)
    ter = _NS(cfg=_NS(terrain_type=terrain_type, terrain_generator=gen if terrain_type == "generator" else None),
              terrain_types=torch.zeros(num_envs, dtype=torch.long) if terrain_types is None else terrain_types)
    if terrain_origins is not None:
        ter.terrain_origins = terrain_origins
    return ter


def reference_terrain_utils():
    """The unmodified V/mdp/utils.py module (no isaaclab import at run time)."""
    _install_command_context(torch.zeros(7, 1))
    return sys.modules["robot_lab.tasks.manager_based.locomotion.velocity.mdp.utils"]


def reference_command_compute(spec: StepSpec, st: dict, uniform_table: torch.Tensor, terrain_type: str,
                              terrain=None, was_on_pit: torch.Tensor | None = None) -> dict:
    """One CommandTerm.compute(dt) through the reference's UniformThresholdVelocityCommand class. ``terrain`` (a
    ``fake_terrain`` with origins) + ``was_on_pit`` exercise the pit branch of _update_command (:61-85)."""
    commands = _install_command_context(uniform_table)
    env = FakeEnv(spec, st)
    c = spec.command
    env.scene.terrain = terrain if terrain is not None else fake_terrain(spec.layout.terrain, terrain_type, env.num_envs)
    cfg = _NS(asset_name="robot", resampling_time_range=c.resampling_time, rel_standing_envs=c.rel_standing_envs,
              rel_heading_envs=c.rel_heading_envs, heading_command=c.heading_command,
              heading_control_stiffness=c.heading_control_stiffness,
              ranges=_NS(lin_vel_x=c.lin_vel_x, lin_vel_y=c.lin_vel_y, ang_vel_z=c.ang_vel_z, heading=c.heading))
    term = commands.UniformThresholdVelocityCommand(cfg, env)
    term.vel_command_b = st["command"].clone()
    term.heading_target = st["heading_target"].clone()
    term.time_left = st["time_left"].clone()
    term.is_heading_env = st["is_heading_env"].clone()
    term.is_standing_env = st["is_standing_env"].clone()
    term.metrics["error_vel_xy"] = st["metric_error_vel_xy"].clone()
    term.metrics["error_vel_yaw"] = st["metric_error_vel_yaw"].clone()
    if was_on_pit is not None:
        term.was_on_pit = was_on_pit.clone()
    term.compute(spec.step_dt)
    return {
        "was_on_pit": term.was_on_pit,
        "command": term.vel_command_b, "heading_target": term.heading_target, "time_left": term.time_left,
        "is_heading_env": term.is_heading_env, "is_standing_env": term.is_standing_env,
        "metric_error_vel_xy": term.metrics["error_vel_xy"], "metric_error_vel_yaw": term.metrics["error_vel_yaw"],
    }


# ------------------------------------------------------------------------------------------------
# reset event: the reference's reset_root_state_uniform (V/mdp/events.py:205-271) on a fake asset
# ------------------------------------------------------------------------------------------------
def reference_reset_root_state(spec: StepSpec, st: dict, ids: torch.Tensor, cfg, env_origins: torch.Tensor,
                               uniforms: torch.Tensor, terrain=None) -> dict:
    """Run the unmodified ``reset_root_state_uniform`` for ``ids`` with the given uniforms ([12+, N]: pose 6, velocity
    6) and return what it writes into the simulator (root pose + velocity of those envs)."""
    _install_command_context(torch.zeros(7, st["root_quat_w"].shape[0]))   # package context for the relative imports
    import importlib.util

    pkg = "robot_lab.tasks.manager_based.locomotion.velocity.mdp"
    full = f"{pkg}.events"
    if full not in sys.modules:
        sp = importlib.util.spec_from_file_location(full, shim.MDP_DIR / "events.py")
        module = importlib.util.module_from_spec(sp)
        module.__package__ = pkg
        sys.modules[full] = module
        sp.loader.exec_module(module)
    events = sys.modules[full]
    env = FakeEnv(spec, st)
    n = env.num_envs
    asset = env.scene["robot"]
    drs = torch.tensor([0.0, 0.0, spec.layout.asset.init_root_height, 1.0, 0.0, 0.0, 0.0] + [0.0] * 6)
    asset.data.default_root_state = drs.unsqueeze(0).repeat(n, 1)
    pose_w = torch.full((n, 7), float("nan"))
    vel_w = torch.full((n, 6), float("nan"))

    def write_pose(pose, env_ids=None):
        pose_w[env_ids] = pose

    def write_vel(vel, env_ids=None):
        vel_w[env_ids] = vel

    asset.write_root_pose_to_sim, asset.write_root_velocity_to_sim = write_pose, write_vel
    env.scene.env_origins = env_origins
    if terrain is not None:
        env.scene.terrain = terrain
    ids = ids.long()
    math_mod = sys.modules["isaaclab.utils.math"]
    rnd_ids = ids
    if terrain is not None:   # the reference samples only for the non-pit envs (V/mdp/events.py:247-265)
        utils = sys.modules[f"{pkg}.utils"]
        rnd_ids = ids[~utils.is_env_assigned_to_terrain(env, "pits")[ids]]
    math_mod._uniform_queue[:] = [uniforms[0:6, rnd_ids].t().contiguous(), uniforms[6:12, rnd_ids].t().contiguous()]
    events.reset_root_state_uniform(env, ids, dict(cfg.pose_range), dict(cfg.velocity_range))
    assert (not math_mod._uniform_queue) or len(rnd_ids) == 0
    math_mod._uniform_queue[:] = []
    assert not torch.isnan(pose_w[ids]).any() and not torch.isnan(vel_w[ids]).any()
    return {"root_pos_w": pose_w[ids, 0:3], "root_quat_w": pose_w[ids, 3:7],
            "root_lin_vel_w": vel_w[ids, 0:3], "root_ang_vel_w": vel_w[ids, 3:6]}

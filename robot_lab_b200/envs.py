"""``ManagerBasedRLEnv`` + managers + ``RslRlVecEnvWrapper`` with the API surface the reference's scripts consume.

What the reference relies on (scripts/reinforcement_learning/rsl_rl/train.py:177-224,
play.py:175-248, scripts/tools/zero_agent.py:56-70): ``gym.make(id, cfg=env_cfg)`` -> env with ``.unwrapped``,
``.num_envs``, ``.device``, ``.observation_space`` / ``.action_space``, ``.reset()``, ``.step(a) -> (obs_dict{"policy",
"critic"}, rew[N], terminated[N] bool, truncated[N] bool, extras{"log": ...})``, ``.close()``; the wrapper adds
``num_actions``, ``max_episode_length``, ``episode_length_buf`` (get/set), ``get_observations()``, ``step(actions) ->
(obs, rew, dones long, extras with "time_outs")``. Term functions read ``env.scene[...]``, ``env.command_manager``,
``env.action_manager`` - those attributes exist here too, backed by the device buffers.

In IsaacLab each manager loops over its terms in Python. Here ``env.step()`` is three CUDA launches
(``rl_process_action`` -> state provider -> ``rl_step`` -> provider reset -> post-reset launch); the manager
objects are views onto the buffers those launches fill, plus per-manager ``compute()`` entry points that run just
their phase of the kernel. There is no CPU implementation behind any of this.

Physics is not part of this tier: a ``StateProvider`` fills the state buffers each step. ``SyntheticStateProvider``
(the default) draws fresh synthetic state on the device; a simulator integration implements the same two methods.
"""

from __future__ import annotations

import math
from typing import Any

import torch

from . import _native as nat
from .cfg import RewardTermCfg
from .engine import MdpStepEngine
from .spec import StepSpec, compact_layout, compile_reward_term, compile_step_spec
from .state import StateBuffers
from .synthetic import make_state

try:  # optional: real gymnasium spaces / registration when the package exists
    import gymnasium as gym  # type: ignore
except Exception:  # pragma: no cover - not installed in the build image
    gym = None

try:
    from tensordict import TensorDict  # type: ignore
except Exception:  # pragma: no cover
    TensorDict = None


class Box:
    """Minimal stand-in for ``gymnasium.spaces.Box`` (shape / dtype / bounds) when gymnasium is absent."""

    def __init__(self, low: float, high: float, shape: tuple[int, ...], dtype=torch.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape})"


# --------------------------------------------------------------------------------------------------
# state providers (the physics / sensor side of the boundary)
# --------------------------------------------------------------------------------------------------
class StateProvider:
    """Producer of the per-step state buffers (root / joint state, contact history, timers, feet kinematics, rays)."""

    def advance(self, env: "ManagerBasedRLEnv") -> None:  # after process_action, before the fused step
        raise NotImplementedError

    def reset(self, env: "ManagerBasedRLEnv", env_ids: torch.Tensor, n_ids: torch.Tensor) -> None:
        """Write post-reset physical state for ``env_ids[: n_ids]`` (device tensors; no host sync required)."""


class SyntheticStateProvider(StateProvider):
    """Rotates over a few pre-generated synthetic state sets (device copies) - the benchmark's producer."""

    def __init__(self, spec: StepSpec, num_envs: int, device, num_sets: int = 4, seed: int = 1234, rank: int = 0):
        self.sets = []
        for i in range(num_sets):
            st = make_state(spec, num_envs, seed=seed + 1000 * i, rank=rank)
            self.sets.append({k: v for k, v in st.items() if k in nat._STATE_FIELDS})
        self.i = 0
        self._staged = None

    def advance(self, env):
        st = self.sets[self.i % len(self.sets)]
        self.i += 1
        for name, val in st.items():
            env.buffers._to_device(name, val)

    def reset(self, env, env_ids, n_ids):
        return  # synthetic state has no notion of a physical reset


class ResetEventStateProvider(SyntheticStateProvider):
    """Synthetic producer whose ``reset`` applies the reference's mode="reset" state events on the device
    (``reset_root_state_uniform`` V/mdp/events.py:205-271 + ``reset_joints_by_scale`` [IL]) through
    ``rl_reset_scene_state`` - the post-reset root / joint state the observation launch then reads."""

    def __init__(self, spec: StepSpec, num_envs: int, device, reset_cfg=None, env_origins: torch.Tensor | None = None, **kw):
        super().__init__(spec, num_envs, device, **kw)
        from .cfg import ResetStateCfg

        self.reset_cfg = reset_cfg if reset_cfg is not None else ResetStateCfg()
        self.env_origins = env_origins if env_origins is not None else torch.zeros(num_envs, 3, device=device)

    def reset(self, env, env_ids, n_ids):
        env.engine.reset_scene_state(env.buffers, self.reset_cfg, self.env_origins, env_ids=env_ids, n_env_ids=n_ids,
                                     seed=env.seed, env_id_offset=env.rank * env.num_envs, use_step_counter=True)


class SensorChainStateProvider(ResetEventStateProvider):
    """Synthetic root / joint / contact-force state, with everything the library can derive from it computed on the
    device in IsaacLab's order for every physics sub-step (``decimation`` of them, V/velocity_env_cfg.py:714):
    ``rl_actuator_step`` (joint target -> applied torque), ``rl_contact_sensor_update`` (force
    history + air / contact timers, row 1) and - once per env step, on the final pose - ``rl_height_scan_cast`` (the
    height scanner over a height field, row 4). ``reset`` is ``rl_reset_scene_state`` (row 2)."""

    def __init__(self, spec: StepSpec, num_envs: int, device, height_field=None, decimation: int = 4,
                 physics_dt: float = 0.005, **kw):
        super().__init__(spec, num_envs, device, **kw)
        self.height_field, self.decimation, self.physics_dt = height_field, decimation, physics_dt
        self.forces = [st["net_forces_w_history"][:, 0].contiguous().to(device) for st in self.sets]
        derived = ("applied_torque", "net_forces_w_history", "current_air_time", "last_air_time", "current_contact_time",
                   "last_contact_time") + (("ray_hits_z", "ray_sensor_pos_z") if height_field is not None else ())
        for st in self.sets:
            for k in derived:
                st.pop(k, None)
        self.substep = 0

    def advance(self, env):
        i = self.i % len(self.sets)
        super().advance(env)
        b, eng = env.buffers, env.engine
        # ring-slot writes move no data, but feet_stumble reads history slot 0 as "the newest sample" (net_forces_w,
        # V/mdp/rewards.py:431-434): a spec with that term active keeps IsaacLab's rolling history
        ring = not any(r.type_name == "feet_stumble" and r.weight != 0.0 for r in env.spec.rewards)
        for _ in range(self.decimation):
            eng.actuator_step(b)
            eng.contact_sensor_update(b, self.forces[i], self.physics_dt, ring_slot=(self.substep % env.spec.T) if ring else -1)
            self.substep += 1
        if self.height_field is not None and env.spec.R > 0:
            eng.height_scan_cast(b, self.height_field)


class ReplayStateProvider(StateProvider):
    """Feeds a fixed list of logical state dicts, one per step (tests: the oracle replays the same list)."""

    def __init__(self, states: list[dict]):
        self.states, self.i = states, 0

    def advance(self, env):
        st = self.states[self.i]
        self.i += 1
        for name in nat._STATE_FIELDS:
            env.buffers._to_device(name, st[name])


# --------------------------------------------------------------------------------------------------
# scene views (env.scene["robot"].data.*, env.scene.sensors["contact_forces"].data.*)
# --------------------------------------------------------------------------------------------------
class _BufferView:
    def __init__(self, env, mapping: dict[str, str]):
        object.__setattr__(self, "_env", env)
        object.__setattr__(self, "_map", mapping)

    def __getattr__(self, name):
        m = object.__getattribute__(self, "_map")
        if name in m:
            return object.__getattribute__(self, "_env").buffers.logical(m[name])
        raise AttributeError(name)


class _Articulation:
    def __init__(self, env):
        self._env = env
        names = ("root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w", "joint_pos", "joint_vel", "joint_acc",
                 "applied_torque", "body_pos_w", "body_lin_vel_w")
        mapping = {n: n for n in names}
        mapping.update(root_link_pos_w="root_pos_w", root_link_quat_w="root_quat_w", body_link_pos_w="body_pos_w")
        self.data = _BufferView(env, mapping)
        asset = env.spec.layout.asset
        self.joint_names, self.body_names = list(asset.joint_names), list(env.spec.layout.asset_body_names)
        self.num_joints = asset.num_joints

    def find_joints(self, name_keys, preserve_order: bool = False):
        from .cfg import resolve_matching_names

        return resolve_matching_names(name_keys, self.joint_names, preserve_order)


class _ContactSensor:
    def __init__(self, env):
        self._env = env
        names = ("net_forces_w_history", "current_air_time", "last_air_time", "current_contact_time", "last_contact_time")
        self.data = _BufferView(env, {n: n for n in names})
        self.body_names = list(env.spec.layout.time_body_names)

    def find_bodies(self, name_keys, preserve_order: bool = False):
        from .cfg import resolve_matching_names

        return resolve_matching_names(name_keys, self.body_names, preserve_order)

    def compute_first_contact(self, dt: float, abs_tol: float = 1.0e-8) -> torch.Tensor:
        t = self.data.current_contact_time
        return (t > 0.0) & (t < (dt + abs_tol))


class _Scene(dict):
    def __init__(self, env):
        super().__init__(robot=_Articulation(env))
        self.sensors = {"contact_forces": _ContactSensor(env)}
        self["contact_forces"] = self.sensors["contact_forces"]
        self.num_envs = env.num_envs
        self.cfg = env.cfg.scene


# --------------------------------------------------------------------------------------------------
# managers
# --------------------------------------------------------------------------------------------------
class _ManagerBase:
    def __init__(self, env: "ManagerBasedRLEnv"):
        self._env = env

    @property
    def num_envs(self):
        return self._env.num_envs

    @property
    def device(self):
        return self._env.device


class ActionManager(_ManagerBase):
    @property
    def action(self) -> torch.Tensor:
        return self._env.buffers.logical("action")

    @property
    def prev_action(self) -> torch.Tensor:
        return self._env.buffers.logical("prev_action")

    @property
    def total_action_dim(self) -> int:
        return self._env.spec.A

    @property
    def processed_actions(self) -> torch.Tensor:
        """Joint position targets in native joint order (what ``apply_action`` would hand to the articulation)."""
        return self._env.buffers.logical("joint_target")

    def process_action(self, action: torch.Tensor) -> None:
        env = self._env
        if action.shape != (env.num_envs, env.spec.A):
            raise ValueError(f"Invalid action shape, expected: {(env.num_envs, env.spec.A)}, received: {tuple(action.shape)}.")
        env.buffers.t["new_action"].copy_(action.to(env.device, torch.float32))
        env.engine.process_action(env.buffers)


class RewardManager(_ManagerBase):
    def __init__(self, env):
        super().__init__(env)
        self.active_terms = [t.name for t in env.spec.rewards]
        self._term_cfgs = dict(env.cfg.rewards.active(RewardTermCfg))

    @property
    def _episode_sums(self) -> dict[str, torch.Tensor]:
        sums = self._env.buffers.logical("episode_sums")
        return {name: sums[:, k] for k, name in enumerate(self.active_terms)}

    @property
    def _step_reward(self) -> torch.Tensor:
        return self._env.buffers.logical("step_reward")

    def get_term_cfg(self, name: str) -> RewardTermCfg:
        return self._term_cfgs[name]

    def compute(self, dt: float | None = None) -> torch.Tensor:
        """Reward phase only (the fused ``env.step()`` normally covers it)."""
        env = self._env
        env.engine.step(env.buffers, phases=nat.PHASE_REWARDS, **env._rng_kwargs())
        return env.buffers.reward

    def evaluate_term(self, func, params: dict[str, Any]) -> torch.Tensor:
        """``func(env, **params)``: one reward term alone on the GPU (``rl_term_eval``)."""
        env = self._env
        term = compile_reward_term(func.__name__, RewardTermCfg(func=func, weight=1.0, params=params), env.spec.layout)
        return env.engine.term_eval(term, env.buffers, terminated=env.buffers.terminated)


class TerminationManager(_ManagerBase):
    def __init__(self, env):
        super().__init__(env)
        self.active_terms = [t.name for t in env.spec.dones]

    @property
    def terminated(self) -> torch.Tensor:
        return self._env.buffers.terminated.bool()

    @property
    def time_outs(self) -> torch.Tensor:
        return self._env.buffers.truncated.bool()

    @property
    def dones(self) -> torch.Tensor:
        return self.terminated | self.time_outs

    def get_term(self, name: str) -> torch.Tensor:
        return ((self._env.buffers.done_bits >> self.active_terms.index(name)) & 1).bool()

    def evaluate_term(self, func, params):
        raise NotImplementedError("termination terms are evaluated inside the fused step; read get_term(name)")


class CommandManager(_ManagerBase):
    def get_command(self, name: str) -> torch.Tensor:
        if name != "base_velocity":
            raise KeyError(name)
        return self._env.buffers.logical("command")

    def get_term(self, name: str):
        if name != "base_velocity":
            raise KeyError(name)
        return self._env.cfg.commands.base_velocity and _CommandTermView(self._env)

    def compute(self, dt: float | None = None) -> None:
        env = self._env
        env.engine.step(env.buffers, phases=nat.PHASE_COMMAND, **env._rng_kwargs())


class _CommandTermView:
    def __init__(self, env):
        self._env, self.cfg = env, env.cfg.commands.base_velocity

    @property
    def command(self):
        return self._env.buffers.logical("command")

    @property
    def metrics(self):
        b = self._env.buffers
        return {"error_vel_xy": b.logical("metric_error_vel_xy"), "error_vel_yaw": b.logical("metric_error_vel_yaw")}

    def __getattr__(self, name):
        if name in ("heading_target", "time_left", "is_heading_env", "is_standing_env"):
            return self._env.buffers.logical(name)
        raise AttributeError(name)


class ObservationManager(_ManagerBase):
    @property
    def group_obs_dim(self) -> dict[str, tuple[int, ...]]:
        return {g.name: (g.dim,) for g in self._env.spec.obs if g.dim > 0}

    def compute(self) -> dict[str, torch.Tensor]:
        env = self._env
        env.engine.step(env.buffers, phases=nat.PHASE_OBS, **env._rng_kwargs())
        return env._obs_dict()

    def evaluate_term(self, func, params):
        raise NotImplementedError("observation terms are evaluated inside the fused step; slice the group row instead")


# --------------------------------------------------------------------------------------------------
# the env
# --------------------------------------------------------------------------------------------------
class ManagerBasedRLEnv:
    """Drop-in for ``isaaclab.envs.ManagerBasedRLEnv`` on the MDP side (no simulator attached)."""

    metadata = {"render_modes": [None]}

    def __init__(self, cfg, render_mode: str | None = None, state_provider: StateProvider | None = None,
                 body_tensors: str | None = None, rank: int = 0, **kwargs):
        self.cfg = cfg
        self.render_mode = render_mode
        self.device = torch.device(cfg.sim.device)
        self.num_envs = int(cfg.scene.num_envs)
        mode = body_tensors or cfg.scene.body_tensors
        layout = compact_layout(cfg) if mode == "compact" else cfg.scene.make_layout()
        self.spec = compile_step_spec(cfg, layout)
        self.engine = MdpStepEngine(self.spec, self.device)  # raises without the CUDA library: no CPU path
        self.buffers: StateBuffers = self.engine.new_buffers(self.num_envs)
        self.seed = int(getattr(cfg, "seed", 0) or 0)
        self.rank = rank
        self.step_dt = self.spec.step_dt
        self.physics_dt = cfg.sim.dt
        self.max_episode_length_s = cfg.episode_length_s
        self.max_episode_length = self.spec.max_episode_length
        self.common_step_counter = 0
        self.extras: dict[str, Any] = {}
        self.scene = _Scene(self)
        self.action_manager = ActionManager(self)
        self.observation_manager = ObservationManager(self)
        self.reward_manager = RewardManager(self)
        self.termination_manager = TerminationManager(self)
        self.command_manager = CommandManager(self)
        self.state_provider = state_provider or SyntheticStateProvider(self.spec, self.num_envs, self.device, seed=1234, rank=rank)
        self.single_observation_space = {g.name: Box(-math.inf, math.inf, (g.dim,)) for g in self.spec.obs if g.dim > 0}
        self.observation_space = {g.name: Box(-math.inf, math.inf, (self.num_envs, g.dim)) for g in self.spec.obs if g.dim > 0}
        self.single_action_space = Box(-math.inf, math.inf, (self.spec.A,))
        self.action_space = Box(-math.inf, math.inf, (self.num_envs, self.spec.A))
        # terrain-aware command restriction (V/mdp/commands.py:61-85): only when the terrain has a "pits" sub-terrain
        self.pit_grid = None
        from . import terrain as _terrain

        if _terrain.terrain_column_range(layout.terrain, "pits") is not None:
            self.pit_grid = _terrain.TerrainGridBuffers.create(layout.terrain, "pits", self.device,
                                                               origins=kwargs.get("terrain_origins"))
            self.was_on_pit = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        self._all_ids = torch.arange(self.num_envs, dtype=torch.int32, device=self.device)
        self._n_all = torch.tensor([self.num_envs], dtype=torch.int32, device=self.device)
        self._closed = False

    # -- properties the scripts / wrapper read --------------------------------------------------------
    @property
    def unwrapped(self):
        return self

    @property
    def episode_length_buf(self) -> torch.Tensor:
        return self.buffers.t["episode_length"].long()

    @episode_length_buf.setter
    def episode_length_buf(self, value: torch.Tensor) -> None:
        self.buffers.t["episode_length"].copy_(value.to(self.device).to(torch.int32))

    @property
    def reset_buf(self) -> torch.Tensor:
        return self.termination_manager.dones

    @property
    def reward_buf(self) -> torch.Tensor:
        return self.buffers.reward

    def _rng_kwargs(self) -> dict:
        return dict(seed=self.seed, env_id_offset=self.rank * self.num_envs, use_random_inputs=False, use_step_counter=True)

    def _obs_dict(self) -> dict[str, torch.Tensor]:
        return {g.name: self.buffers.obs[i] for i, g in enumerate(self.spec.obs) if g.dim > 0}

    def _log(self) -> dict[str, torch.Tensor]:
        b, log = self.buffers, {}
        for k, name in enumerate(self.reward_manager.active_terms):
            log[f"Episode_Reward/{name}"] = b.log_episode_sum_mean[k] / self.max_episode_length_s
        for i, name in enumerate(self.termination_manager.active_terms):
            log[f"Episode_Termination/{name}"] = b.log_done_term_count[i]
        log["Metrics/base_velocity/error_vel_xy"] = b.log_metric_mean[0]
        log["Metrics/base_velocity/error_vel_yaw"] = b.log_metric_mean[1]
        return log

    # -- gym API -------------------------------------------------------------------------------------------
    def reset(self, seed: int | None = None, options: dict | None = None):
        """Reset every env: provider state, manager reset of all ids, command compute, observations."""
        if seed is not None:
            self.seed = int(seed)
        b = self.buffers
        self.state_provider.advance(self)
        self.state_provider.reset(self, self._all_ids, self._n_all)
        b.done_bits.zero_()
        self.engine.step(b, phases=nat.PHASE_RESET | nat.PHASE_COMMAND | nat.PHASE_OBS, env_ids=self._all_ids,
                         n_env_ids=self._n_all, **self._rng_kwargs())
        self.extras = {"log": self._log()}
        return self._obs_dict(), self.extras

    def step(self, action: torch.Tensor):
        """ManagerBasedRLEnv.step() [IL] order: three launches of this library + the provider."""
        b, eng = self.buffers, self.engine
        self.action_manager.process_action(action)                                   # 1 (+ common step counter)
        self.state_provider.advance(self)                                            # 2 physics / sensors
        self.common_step_counter += 1
        eng.step_pre_reset(b, **self._rng_kwargs())                                  # 3-5: dones, rewards, reset ids
        self.state_provider.reset(self, b.reset_ids, b.n_reset)                      # 6a external reset
        if self.pit_grid is None:
            eng.step_post_reset(b, **self._rng_kwargs())                             # 6b manager reset, 7 command, 9 obs
        else:   # the pit branch of _update_command sits between the command update and the observations
            rng = self._rng_kwargs()
            eng.step(b, phases=nat.PHASE_RESET | nat.PHASE_COMMAND, **rng)
            eng.command_pit_restrict(b, self.pit_grid, self.was_on_pit, **rng)
            eng.step(b, phases=nat.PHASE_OBS, **rng)
        self.extras = {"log": self._log()}
        return self._obs_dict(), b.reward, b.terminated.bool(), b.truncated.bool(), self.extras

    def close(self) -> None:
        if not self._closed:
            self.engine.close()
            self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RslRlVecEnvWrapper:
    """``isaaclab_rl.rsl_rl.RslRlVecEnvWrapper`` [IL] surface: what rsl_rl's OnPolicyRunner calls."""

    def __init__(self, env: ManagerBasedRLEnv, clip_actions: float | None = None):
        if not isinstance(env.unwrapped, ManagerBasedRLEnv):
            raise ValueError(f"The environment must be inherited from ManagerBasedRLEnv. Environment type: {type(env)}")
        self.env = env
        self.clip_actions = clip_actions
        self.num_envs = env.unwrapped.num_envs
        self.device = env.unwrapped.device
        self.max_episode_length = env.unwrapped.max_episode_length
        self.num_actions = env.unwrapped.action_manager.total_action_dim
        self.env.reset()

    def __str__(self):
        return f"<{type(self).__name__}{self.env}>"

    @classmethod
    def class_name(cls) -> str:
        return cls.__name__

    @property
    def cfg(self):
        return self.unwrapped.cfg

    @property
    def unwrapped(self) -> ManagerBasedRLEnv:
        return self.env.unwrapped

    @property
    def observation_space(self):
        return self.env.observation_space

    @property
    def action_space(self):
        return self.env.action_space

    @property
    def episode_length_buf(self) -> torch.Tensor:
        return self.unwrapped.episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, value: torch.Tensor):
        self.unwrapped.episode_length_buf = value

    def seed(self, seed: int = -1) -> int:
        self.unwrapped.seed = seed
        return seed

    def _pack(self, obs: dict[str, torch.Tensor]):
        if TensorDict is not None:
            return TensorDict(obs, batch_size=[self.num_envs])
        return obs

    def reset(self):
        obs, extras = self.env.reset()
        return self._pack(obs), extras

    def get_observations(self):
        return self._pack(self.unwrapped._obs_dict())

    def step(self, actions: torch.Tensor):
        if self.clip_actions is not None:
            actions = torch.clamp(actions, -self.clip_actions, self.clip_actions)
        obs, rew, terminated, truncated, extras = self.env.step(actions)
        dones = (terminated | truncated).to(dtype=torch.long)
        if not getattr(self.unwrapped.cfg, "is_finite_horizon", False):
            extras["time_outs"] = truncated
        return self._pack(obs), rew, dones, extras

    def close(self):
        return self.env.close()


def make(task: str, cfg=None, num_envs: int | None = None, device: str | None = None, **kwargs) -> ManagerBasedRLEnv:
    """``gym.make(task, cfg=env_cfg)`` equivalent (V/config/quadruped/unitree_go2/__init__.py:12-32)."""
    from .tasks import make_env_cfg

    if cfg is None:
        cfg = make_env_cfg(task, num_envs)
    elif num_envs is not None:
        cfg.scene.num_envs = int(num_envs)
    if device is not None:
        cfg.sim.device = device
    return ManagerBasedRLEnv(cfg, **kwargs)


if gym is not None:  # pragma: no cover - registration under the reference's ids when gymnasium exists
    from .tasks import TASKS as _TASKS

    for _task in _TASKS:
        try:
            gym.register(id=_task, entry_point="robot_lab_b200.envs:ManagerBasedRLEnv", disable_env_checker=True,
                         kwargs={"env_cfg_entry_point": f"robot_lab_b200.tasks:TASKS['{_task}']"})
        except Exception:
            pass

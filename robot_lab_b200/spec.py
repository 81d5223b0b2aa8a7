"""Spec compiler: task cfg tree -> flat, fully resolved ``StepSpec`` -> ``RlStepSpec`` (C-ABI POD).

This does at construction time what the reference's managers do when they are built (IsaacLab manager
``_prepare_terms`` [IL]): resolve every ``SceneEntityCfg`` regex to joint / body indices, expand per-joint
dicts (action scale / clip, GO2/rough_env_cfg.py:51-53), fix the term order and collect params. The result is
a plain-Python description (consumed as-is by the CPU oracle in ``oracle/``) plus its ctypes image (uploaded to
``__constant__`` memory by ``rl_ctx_create``).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Sequence

from . import _native as nat
from .assets import RobotAsset
from .cfg import (
    JointPositionActionCfg,
    JointVelocityActionCfg,
    ObservationGroupCfg,
    ObservationTermCfg,
    RewardTermCfg,
    SceneEntityCfg,
    TerminationTermCfg,
    TerrainCfg,
    UniformThresholdVelocityCommandCfg,
    resolve_matching_names,
    resolve_matching_names_values,
)


@dataclass
class SceneLayout:
    """Which bodies each body-indexed state tensor carries (its "index space").

    In IsaacLab the contact sensor and the articulation expose *all* bodies; a B200-native provider only has to
    materialise the bodies some term reads. Terms are resolved against the full body list first
    (reference semantics) and then mapped into these spaces by name.
    """

    asset: RobotAsset
    hist_body_names: tuple[str, ...]   # bodies in net_forces_w_history [.., T, B, 3]
    time_body_names: tuple[str, ...]   # bodies in current/last air/contact time
    asset_body_names: tuple[str, ...]  # bodies in body_pos_w / body_lin_vel_w
    terrain: TerrainCfg
    num_rays: int = 0
    hist_len: int = 3

    @staticmethod
    def full(asset: RobotAsset, terrain: TerrainCfg, num_rays: int, hist_len: int = 3) -> "SceneLayout":
        """IsaacLab-shaped tensors: every body everywhere."""
        return SceneLayout(asset, asset.body_names, asset.body_names, asset.body_names, terrain, num_rays, hist_len)


@dataclass
class RewardTermSpec:
    name: str
    type_name: str
    type_id: int
    weight: float
    p: list[float] = field(default_factory=lambda: [0.0] * 6)
    joint_ids: list[int] = field(default_factory=list)
    body_ids: list[int] = field(default_factory=list)  # history-body space
    idx_a: list[int] = field(default_factory=list)
    idx_b: list[int] = field(default_factory=list)
    idx_c: list[int] = field(default_factory=list)
    n_idx: int = 0


@dataclass
class ObsTermSpec:
    name: str
    type_name: str
    type_id: int
    dim: int
    noise: tuple[float, float] | None = None
    clip: tuple[float, float] | None = None
    scale: float | None = None
    p: list[float] = field(default_factory=lambda: [0.0, 0.0])
    ids: list[int] = field(default_factory=list)
    zero_cols: list[int] = field(default_factory=list)


@dataclass
class ObsGroupSpec:
    name: str
    enable_corruption: bool
    terms: list[ObsTermSpec]

    @property
    def dim(self) -> int:
        return sum(t.dim for t in self.terms)


@dataclass
class DoneTermSpec:
    name: str
    type_name: str
    type_id: int
    time_out: bool
    p: list[float] = field(default_factory=lambda: [0.0] * 4)
    body_ids: list[int] = field(default_factory=list)


@dataclass
class CommandSpec:
    resampling_time: tuple[float, float]
    rel_standing_envs: float
    rel_heading_envs: float
    heading_command: bool
    heading_control_stiffness: float
    lin_vel_x: tuple[float, float]
    lin_vel_y: tuple[float, float]
    ang_vel_z: tuple[float, float]
    heading: tuple[float, float]
    small_cmd_threshold: float
    max_command_step: float


@dataclass
class ActionSpec:
    joint_ids: list[int]
    joint_names: list[str]
    scale: list[float]
    offset: list[float]
    clip: list[tuple[float, float]] | None
    kind: list[int] | None = None   # per column: 0 joint position target, 1 joint velocity target (None = all 0)


@dataclass
class StepSpec:
    task: str
    layout: SceneLayout
    step_dt: float
    max_episode_length: int
    max_episode_length_s: float
    contact_time_abs_tol: float
    default_joint_pos: list[float]
    default_joint_vel: list[float]
    soft_pos_limits: list[tuple[float, float]]
    soft_vel_limits: list[float]
    rewards: list[RewardTermSpec]
    dones: list[DoneTermSpec]
    obs: list[ObsGroupSpec]  # [policy, critic]
    command: CommandSpec
    action: ActionSpec

    # ---- dims ----
    @property
    def J(self) -> int:
        return self.layout.asset.num_joints

    @property
    def B(self) -> int:
        return len(self.layout.hist_body_names)

    @property
    def T(self) -> int:
        return self.layout.hist_len

    @property
    def Bt(self) -> int:
        return len(self.layout.time_body_names)

    @property
    def Ba(self) -> int:
        return len(self.layout.asset_body_names)

    @property
    def R(self) -> int:
        return self.layout.num_rays

    @property
    def A(self) -> int:
        return len(self.action.joint_ids)

    @property
    def K(self) -> int:
        return len(self.rewards)

    def algorithmic_bytes_per_env_step(self, bodies_read: int | None = None) -> int:
        """SURVEY.md 8(d) byte formula (fp32 words read + written per env-step)."""
        J, F, R, K = self.J, self.Bt, self.R, self.K
        B = self.B if bodies_read is None else bodies_read
        pol, crit = self.obs[0].dim, self.obs[1].dim
        read = 13 + 6 * J + 3 + 9 * B + 4 * F + 6 * F + R + 1 + K
        write = pol + crit + 1 + 2 * K + 2 * J
        return 4 * (read + write) + 3

    def algorithmic_bytes_per_launch(self, kind: str) -> int:
        """The same accounting split over the two launches of an env step. ``pre_reset`` = DONES |
        REWARDS | COMPACT, ``post_reset`` = RESET | COMMAND | OBS over all envs. Root state, joint state and the
        command are read by both, so the two add up to slightly more than the fused figure."""
        J, F, R, K, B = self.J, self.Bt, self.R, self.K, self.B
        pol, crit = self.obs[0].dim, self.obs[1].dim
        if kind == "pre_reset":
            read = 13 + 6 * J + 3 + 9 * B + 4 * F + 6 * F + K + 1      # + episode length
            write = 1 + 2 * K + 1
            return 4 * (read + write) + 3
        if kind == "post_reset":
            read = 13 + 3 * J + 3 + 4 + R + 1                          # + heading target, timer, 2 metrics
            write = pol + crit + 3 + 4
            return 4 * (read + write) + 2                              # + the two command flag bytes
        raise ValueError(kind)

    # ---- ctypes image ----
    def to_ctypes(self) -> nat.RlStepSpec:
        s = nat.RlStepSpec()
        s.abi_version = nat.RL_ABI_VERSION
        s.num_joints, s.num_hist_bodies, s.hist_len = self.J, self.B, self.T
        s.num_time_bodies, s.num_asset_bodies, s.num_rays = self.Bt, self.Ba, self.R
        s.num_reward_terms, s.num_done_terms = self.K, len(self.dones)
        s.max_episode_length = self.max_episode_length
        s.step_dt = self.step_dt
        s.contact_time_abs_tol = self.contact_time_abs_tol
        _limit("joints", self.J, nat.RL_MAX_JOINTS)
        _limit("history bodies", self.B, nat.RL_MAX_BODIES)
        _limit("time bodies", self.Bt, nat.RL_MAX_TIME_BODIES)
        _limit("asset bodies", self.Ba, nat.RL_MAX_ASSET_BODIES)
        _limit("reward terms", self.K, nat.RL_MAX_REWARD_TERMS)
        _limit("done terms", len(self.dones), nat.RL_MAX_DONE_TERMS)
        for j in range(self.J):
            s.default_joint_pos[j] = self.default_joint_pos[j]
            s.default_joint_vel[j] = self.default_joint_vel[j]
            s.soft_pos_limit_lo[j], s.soft_pos_limit_hi[j] = self.soft_pos_limits[j]
            s.soft_vel_limit[j] = self.soft_vel_limits[j]
        for k, t in enumerate(self.rewards):
            s.rewards[k] = reward_term_to_ctypes(t)
        for k, d in enumerate(self.dones):
            c = s.dones[k]
            c.type, c.time_out = d.type_id, int(d.time_out)
            for i, v in enumerate(d.p):
                c.p[i] = v
            c.body_mask = _mask(d.body_ids)
        for g, grp in enumerate(self.obs):
            cg = s.obs[g]
            _limit(f"obs terms in group {grp.name}", len(grp.terms), nat.RL_MAX_OBS_TERMS)
            cg.n_terms, cg.dim, cg.enable_corruption = len(grp.terms), grp.dim, int(grp.enable_corruption)
            for i, t in enumerate(grp.terms):
                ct = cg.terms[i]
                ct.type, ct.dim = t.type_id, t.dim
                ct.has_noise = int(t.noise is not None)
                if t.noise is not None:
                    ct.noise_lo, ct.noise_hi = t.noise
                ct.has_clip = int(t.clip is not None)
                if t.clip is not None:
                    ct.clip_lo, ct.clip_hi = t.clip
                ct.has_scale = int(t.scale is not None)
                if t.scale is not None:
                    ct.scale = t.scale
                ct.p[0], ct.p[1] = t.p
                ct.zero_mask = _mask(t.zero_cols)
                _limit(f"ids of obs term {t.name}", len(t.ids), nat.RL_MAX_JOINTS)
                for c_i, jid in enumerate(t.ids):
                    ct.ids[c_i] = jid
        cc, cmd = s.command, self.command
        cc.resampling_time_lo, cc.resampling_time_hi = cmd.resampling_time
        cc.rel_standing_envs, cc.rel_heading_envs = cmd.rel_standing_envs, cmd.rel_heading_envs
        cc.heading_command = int(cmd.heading_command)
        cc.heading_control_stiffness = cmd.heading_control_stiffness
        cc.lin_vel_x_lo, cc.lin_vel_x_hi = cmd.lin_vel_x
        cc.lin_vel_y_lo, cc.lin_vel_y_hi = cmd.lin_vel_y
        cc.ang_vel_z_lo, cc.ang_vel_z_hi = cmd.ang_vel_z
        cc.heading_lo, cc.heading_hi = cmd.heading
        cc.small_cmd_threshold = cmd.small_cmd_threshold
        cc.max_command_step = cmd.max_command_step
        ac, act = s.action, self.action
        _limit("actions", self.A, nat.RL_MAX_JOINTS)
        ac.n_actions = self.A
        ac.has_clip = int(act.clip is not None)
        for a in range(self.A):
            ac.joint_ids[a] = act.joint_ids[a]
            ac.target_kind[a] = act.kind[a] if act.kind is not None else 0
            ac.scale[a] = act.scale[a]
            ac.offset[a] = act.offset[a]
            lo, hi = act.clip[a] if act.clip is not None else (-math.inf, math.inf)
            ac.clip_lo[a], ac.clip_hi[a] = lo, hi
        return s


def _limit(what: str, n: int, cap: int) -> None:
    if n > cap:
        raise ValueError(f"{what}: {n} exceeds the ABI limit {cap}")


def _mask(ids: Sequence[int]) -> int:
    m = 0
    for i in ids:
        if not 0 <= i < 64:
            raise ValueError(f"index {i} does not fit a 64-bit mask")
        m |= 1 << i
    return m


def reward_term_to_ctypes(t: RewardTermSpec) -> nat.RlRewardTerm:
    c = nat.RlRewardTerm()
    c.type, c.weight = t.type_id, t.weight
    for i, v in enumerate(t.p):
        c.p[i] = v
    c.joint_mask = _mask(t.joint_ids)
    c.body_mask = _mask(t.body_ids)
    c.n_idx = t.n_idx
    for name in ("idx_a", "idx_b", "idx_c"):
        vals = getattr(t, name)
        _limit(f"{name} of reward term {t.name}", len(vals), nat.RL_MAX_IDX)
        arr = getattr(c, name)
        for i, v in enumerate(vals):
            arr[i] = v
    return c


# Which index lists of a reward term live in which body space (see include/rl_mdp_step.h).
TIME_SPACE_FIELDS = {
    "feet_air_time": ("idx_a",), "feet_air_time_positive_biped": ("idx_a",), "feet_air_time_variance": ("idx_a",),
    "feet_contact": ("idx_a",), "feet_contact_without_cmd": ("idx_a",), "feet_gait": ("idx_a",),
    "wheel_vel_penalty": ("idx_a",),
}
HIST_SPACE_FIELDS = {
    "undesired_contacts": ("body_ids",), "contact_forces": ("body_ids",), "feet_stumble": ("idx_c",),
    "feet_slide": ("idx_c",),
}
ASSET_SPACE_FIELDS = {
    "feet_slide": ("idx_b",), "feet_height": ("idx_b",), "feet_height_body": ("idx_b",),
    "feet_distance_y_exp": ("idx_b",), "feet_distance_xy_exp": ("idx_b",),
}


def compact_layout(env_cfg: Any) -> SceneLayout:
    """Layout carrying only the bodies some active term reads (the B200-native state provider's layout).

    Resolved by a dry compile against the full (IsaacLab-shaped) layout; each space keeps the full list's
    order, so regex-derived body orders are unchanged.
    """
    import copy

    full = env_cfg.scene.make_layout()
    spec = compile_step_spec(copy.deepcopy(env_cfg), full)
    used = {"time": set(), "hist": set(), "asset": set()}
    for t in spec.rewards:
        for space, table in (("time", TIME_SPACE_FIELDS), ("hist", HIST_SPACE_FIELDS), ("asset", ASSET_SPACE_FIELDS)):
            for fld in table.get(t.type_name, ()):
                used[space].update(full.asset.body_names[i] for i in getattr(t, fld))
    for d in spec.dones:
        used["hist"].update(full.asset.body_names[i] for i in d.body_ids)
    pick = lambda s: tuple(n for n in full.asset.body_names if n in used[s])  # noqa: E731
    return SceneLayout(full.asset, pick("hist"), pick("time"), pick("asset"), full.terrain, full.num_rays, full.hist_len)


# ------------------------------------------------------------------------------------------------
# resolution helpers
# ------------------------------------------------------------------------------------------------
def _joint_ids(cfg: SceneEntityCfg | None, asset: RobotAsset) -> list[int]:
    if cfg is None:
        return list(range(asset.num_joints))
    return cfg.resolve_joints(asset.joint_names)


def _body_names(cfg: SceneEntityCfg, asset: RobotAsset) -> list[str]:
    """Names selected from the *full* body list, in the order IsaacLab would return them."""
    ids = cfg.resolve_bodies(asset.body_names)
    return [asset.body_names[i] for i in ids]


def _to_space(names: Sequence[str], space: Sequence[str], what: str, term: str) -> list[int]:
    out = []
    for n in names:
        if n not in space:
            raise ValueError(f"term '{term}' reads body '{n}' which the {what} tensor does not carry ({list(space)})")
        out.append(list(space).index(n))
    return out


def _find(names_re: str | Sequence[str], pool: Sequence[str]) -> list[int]:
    return resolve_matching_names(names_re, pool, preserve_order=False)[0]


def compile_reward_term(name: str, cfg: RewardTermCfg, layout: SceneLayout) -> RewardTermSpec:
    """RewardTermCfg(func, weight, params) -> resolved term (index spaces per include/rl_mdp_step.h)."""
    func = cfg.func
    if getattr(func, "rl_kind", None) != "reward":
        raise TypeError(f"reward term '{name}': {func!r} is not a robot_lab_b200.mdp reward function")
    asset = layout.asset
    P = dict(func.rl_defaults)
    P.update(cfg.params)
    t = RewardTermSpec(name=name, type_name=func.rl_type_name, type_id=func.rl_type_id, weight=float(cfg.weight))
    ty = func.rl_type_name
    sensor_cfg: SceneEntityCfg | None = P.get("sensor_cfg")
    asset_cfg: SceneEntityCfg | None = P.get("asset_cfg")

    def feet_time():
        return _to_space(_body_names(sensor_cfg, asset), layout.time_body_names, "air/contact-time", name)

    def feet_hist():
        return _to_space(_body_names(sensor_cfg, asset), layout.hist_body_names, "contact-history", name)

    def feet_asset():
        return _to_space(_body_names(asset_cfg, asset), layout.asset_body_names, "body pos/vel", name)

    if ty in ("is_terminated", "lin_vel_z_l2", "ang_vel_xy_l2", "flat_orientation_l2", "upward", "action_rate_l2"):
        pass
    elif ty == "base_height_l2":
        if P.get("sensor_cfg") is not None:
            raise NotImplementedError(
                "base_height_l2 with a ray sensor has a batch-global fallback branch (V/mdp/rewards.py:634) that "
                "breaks env independence; only the sensor_cfg=None form is supported")
        t.p[0] = float(P["target_height"])
    elif ty in ("joint_torques_l2", "joint_vel_l2", "joint_acc_l2", "joint_deviation_l1", "joint_pos_limits", "joint_power"):
        t.joint_ids = _joint_ids(asset_cfg, asset)
    elif ty == "joint_vel_limits":
        t.joint_ids = _joint_ids(asset_cfg, asset)
        t.p[0] = float(P["soft_ratio"])
    elif ty == "stand_still":
        t.joint_ids = _joint_ids(asset_cfg, asset)
        t.p[0] = float(P["command_threshold"])
    elif ty == "joint_pos_penalty":
        t.joint_ids = _joint_ids(asset_cfg, asset)
        t.p[0], t.p[1], t.p[2] = float(P["stand_still_scale"]), float(P["velocity_threshold"]), float(P["command_threshold"])
    elif ty in ("joint_mirror", "action_mirror"):
        pairs = P["mirror_joints"]
        for re_a, re_b in pairs:
            ids_a, ids_b = _find(re_a, asset.joint_names), _find(re_b, asset.joint_names)
            if len(ids_a) != len(ids_b):
                raise ValueError(f"{name}: mirror pair {re_a!r}/{re_b!r} resolves to different sizes")
            t.idx_a += ids_a
            t.idx_b += ids_b
        t.n_idx = len(t.idx_a)
        t.p[0] = (1 / len(pairs)) if len(pairs) > 0 else 0.0
    elif ty == "action_sync":
        groups = P["joint_groups"]
        for grp in groups:
            ids = []
            for jn in grp:
                ids += _find(jn, asset.joint_names)
            t.idx_b.append(len(t.idx_a))
            t.idx_c.append(len(ids))
            t.idx_a += ids
        t.n_idx = len(groups)
        t.p[0] = (1 / len(groups)) if len(groups) > 0 else 0.0
    elif ty in ("undesired_contacts", "contact_forces"):
        t.p[0] = float(P["threshold"])
        t.body_ids = feet_hist()
    elif ty in ("track_lin_vel_xy_exp", "track_ang_vel_z_exp", "track_lin_vel_xy_yaw_frame_exp", "track_ang_vel_z_world_exp"):
        t.p[0] = float(P["std"]) ** 2
    elif ty in ("feet_air_time", "feet_air_time_positive_biped"):
        t.p[0] = float(P["threshold"])
        t.idx_a = feet_time()
        t.n_idx = len(t.idx_a)
    elif ty in ("feet_air_time_variance", "feet_contact_without_cmd"):
        t.idx_a = feet_time()
        t.n_idx = len(t.idx_a)
    elif ty == "feet_contact":
        t.p[0] = float(P["expect_contact_num"])
        t.idx_a = feet_time()
        t.n_idx = len(t.idx_a)
    elif ty == "feet_gait":
        pairs = P["synced_feet_pair_names"]
        if len(pairs) != 2 or len(pairs[0]) != 2 or len(pairs[1]) != 2:
            raise ValueError("This reward only supports gaits with two pairs of synchronized feet, like trotting.")
        names = []
        for pair in pairs:  # contact_sensor.find_bodies(pair)[0] -> ids in sensor order (V/mdp/rewards.py:188-189)
            ids = _find(list(pair), asset.body_names)
            names += [asset.body_names[i] for i in ids]
        t.idx_a = _to_space(names, layout.time_body_names, "air/contact-time", name)
        t.n_idx = 4
        t.p[0], t.p[1] = float(P["std"]), float(P["max_err"]) ** 2
        t.p[2], t.p[3] = float(P["velocity_threshold"]), float(P["command_threshold"])
    elif ty == "feet_stumble":
        t.idx_c = feet_hist()
        t.n_idx = len(t.idx_c)
    elif ty == "feet_slide":
        t.idx_c = feet_hist()
        t.idx_b = feet_asset()
        if len(t.idx_b) != len(t.idx_c):
            raise ValueError(f"{name}: sensor and asset body lists differ in length")
        t.n_idx = len(t.idx_c)
    elif ty in ("feet_height", "feet_height_body"):
        t.p[0], t.p[1] = float(P["target_height"]), float(P["tanh_mult"])
        t.idx_b = feet_asset()
        t.n_idx = len(t.idx_b)
    elif ty == "feet_distance_y_exp":
        t.p[0], t.p[1] = float(P["stance_width"]), float(P["std"]) ** 2
        t.idx_b = feet_asset()
        t.n_idx = len(t.idx_b)
    elif ty == "feet_distance_xy_exp":
        t.p[0], t.p[1], t.p[2] = float(P["stance_width"]), float(P["stance_length"]), float(P["std"]) ** 2
        t.idx_b = feet_asset()
        if len(t.idx_b) != 4:
            raise ValueError(f"{name}: needs exactly 4 feet (V/mdp/rewards.py:478)")
        t.n_idx = 4
    elif ty == "wheel_vel_penalty":
        t.p[0], t.p[1] = float(P["velocity_threshold"]), float(P["command_threshold"])
        t.idx_a = feet_time()
        t.idx_b = _joint_ids(asset_cfg, asset)
        if len(t.idx_a) != len(t.idx_b):
            raise ValueError(f"{name}: wheel bodies and wheel joints differ in number")
        t.n_idx = len(t.idx_a)
    else:
        raise NotImplementedError(ty)
    return t


def compile_obs_term(name: str, cfg: ObservationTermCfg, layout: SceneLayout, n_actions: int) -> ObsTermSpec:
    func = cfg.func
    if getattr(func, "rl_kind", None) != "obs":
        raise TypeError(f"observation term '{name}': {func!r} is not a robot_lab_b200.mdp observation function")
    asset = layout.asset
    P = dict(func.rl_defaults)
    P.update(cfg.params)
    ty = func.rl_type_name
    t = ObsTermSpec(name=name, type_name=ty, type_id=func.rl_type_id, dim=0)
    if ty in ("base_lin_vel", "base_ang_vel", "projected_gravity", "generated_commands"):
        t.dim = 3
    elif ty in ("joint_pos_rel", "joint_vel_rel"):
        t.ids = _joint_ids(P.get("asset_cfg"), asset)
        t.dim = len(t.ids)
    elif ty == "joint_pos_rel_without_wheel":
        t.ids = _joint_ids(P.get("asset_cfg"), asset)
        t.dim = len(t.ids)
        wheel = P.get("wheel_asset_cfg")
        t.zero_cols = _joint_ids(wheel, asset) if wheel is not None else list(range(t.dim))
    elif ty == "last_action":
        t.dim = n_actions
    elif ty == "height_scan":
        if layout.num_rays <= 0:
            raise ValueError(f"{name}: height_scan without a height scanner in the scene")
        t.dim = layout.num_rays
        t.p[0] = float(P["offset"])
    elif ty == "phase":
        t.dim = 2
        t.p[0] = float(P["cycle_time"])
    else:
        raise NotImplementedError(ty)
    if cfg.noise is not None:
        t.noise = (float(cfg.noise.n_min), float(cfg.noise.n_max))
    if cfg.clip is not None:
        t.clip = (float(cfg.clip[0]), float(cfg.clip[1]))
    if cfg.scale is not None:
        t.scale = float(cfg.scale)
    return t


def compile_done_term(name: str, cfg: TerminationTermCfg, layout: SceneLayout) -> DoneTermSpec:
    func = cfg.func
    if getattr(func, "rl_kind", None) != "done":
        raise TypeError(f"termination term '{name}': {func!r} is not a robot_lab_b200.mdp termination function")
    P = dict(func.rl_defaults)
    P.update(cfg.params)
    ty = func.rl_type_name
    d = DoneTermSpec(name=name, type_name=ty, type_id=func.rl_type_id, time_out=bool(cfg.time_out))
    if ty == "time_out":
        pass
    elif ty == "terrain_out_of_bounds":
        ter = layout.terrain
        if ter.terrain_type == "plane":
            d.p = [0.0, 0.0, 0.0, 0.0]
        elif ter.terrain_type == "generator":
            map_w = ter.num_rows * ter.size[0] + 2 * ter.border_width
            map_h = ter.num_cols * ter.size[1] + 2 * ter.border_width
            buf = float(P["distance_buffer"])
            d.p = [0.5 * map_w - buf, 0.5 * map_h - buf, 1.0, 0.0]
        else:
            raise ValueError("Received unsupported terrain type, must be either 'plane' or 'generator'.")
    elif ty == "illegal_contact":
        d.p[0] = float(P["threshold"])
        names = _body_names(P["sensor_cfg"], layout.asset)
        d.body_ids = _to_space(names, layout.hist_body_names, "contact-history", name)
    else:
        raise NotImplementedError(ty)
    return d


def compile_action(cfg: JointPositionActionCfg, asset: RobotAsset) -> ActionSpec:
    ids, names = resolve_matching_names(cfg.joint_names, asset.joint_names, cfg.preserve_order)
    n = len(ids)
    if isinstance(cfg.scale, dict):
        scale = [1.0] * n
        idx, _, vals = resolve_matching_names_values(cfg.scale, names)
        for i, v in zip(idx, vals):
            scale[i] = float(v)
    else:
        scale = [float(cfg.scale)] * n
    is_vel = isinstance(cfg, JointVelocityActionCfg)
    if cfg.use_default_offset:
        dj = asset.default_joint_vel() if is_vel else asset.default_joint_pos()
        offset = [dj[j] for j in ids]
    elif isinstance(cfg.offset, dict):
        offset = [0.0] * n
        idx, _, vals = resolve_matching_names_values(cfg.offset, names)
        for i, v in zip(idx, vals):
            offset[i] = float(v)
    else:
        offset = [float(cfg.offset)] * n
    clip = None
    if cfg.clip is not None:
        clip = [(-math.inf, math.inf)] * n
        idx, _, vals = resolve_matching_names_values(cfg.clip, names)
        for i, v in zip(idx, vals):
            clip[i] = (float(v[0]), float(v[1]))
    return ActionSpec(joint_ids=ids, joint_names=names, scale=scale, offset=offset, clip=clip, kind=[int(is_vel)] * n)


def compile_actions(actions_cfg: Any, asset: RobotAsset) -> ActionSpec:
    """ActionManager [IL]: the active action terms in declaration order, concatenated into one action vector."""
    terms = [compile_action(c, asset) for _n, c in actions_cfg.active(JointPositionActionCfg)]
    if not terms:
        raise ValueError("no action term")
    any_clip = any(t.clip is not None for t in terms)
    clip = None
    if any_clip:
        clip = []
        for t in terms:
            clip += t.clip if t.clip is not None else [(-math.inf, math.inf)] * len(t.joint_ids)
    return ActionSpec(joint_ids=sum((t.joint_ids for t in terms), []), joint_names=sum((t.joint_names for t in terms), []),
                      scale=sum((t.scale for t in terms), []), offset=sum((t.offset for t in terms), []), clip=clip,
                      kind=sum((t.kind for t in terms), []))


def compile_command(cfg: UniformThresholdVelocityCommandCfg, step_dt: float) -> CommandSpec:
    r = cfg.ranges
    if cfg.heading_command and r.heading is None:
        raise ValueError("The velocity command has heading commands active (heading_command=True) but the `ranges.heading` parameter is set to None.")
    return CommandSpec(
        resampling_time=tuple(cfg.resampling_time_range),
        rel_standing_envs=cfg.rel_standing_envs,
        rel_heading_envs=cfg.rel_heading_envs,
        heading_command=cfg.heading_command,
        heading_control_stiffness=cfg.heading_control_stiffness,
        lin_vel_x=tuple(r.lin_vel_x), lin_vel_y=tuple(r.lin_vel_y), ang_vel_z=tuple(r.ang_vel_z),
        heading=tuple(r.heading) if r.heading is not None else (0.0, 0.0),
        small_cmd_threshold=cfg.small_command_threshold,
        max_command_step=cfg.resampling_time_range[1] / step_dt,
    )


def compile_step_spec(env_cfg: Any, layout: SceneLayout | None = None) -> StepSpec:
    """Flatten a ``LocomotionVelocityRoughEnvCfg``-style cfg tree (see ``robot_lab_b200.tasks``)."""
    if layout is None:
        layout = env_cfg.scene.make_layout()
    asset = layout.asset
    step_dt = env_cfg.sim.dt * env_cfg.decimation
    action = compile_actions(env_cfg.actions, asset)
    rewards = [compile_reward_term(n, c, layout) for n, c in env_cfg.rewards.active(RewardTermCfg)]
    dones = [compile_done_term(n, c, layout) for n, c in env_cfg.terminations.active(TerminationTermCfg)]
    groups = []
    for gname in ("policy", "critic"):
        g: ObservationGroupCfg | None = getattr(env_cfg.observations, gname, None)
        if g is None:
            groups.append(ObsGroupSpec(gname, False, []))
            continue
        terms = [compile_obs_term(n, c, layout, len(action.joint_ids)) for n, c in g.active(ObservationTermCfg)]
        if not g.enable_corruption:
            for t in terms:  # noise cfgs are dropped at init when corruption is off [IL]
                t.noise = None
        groups.append(ObsGroupSpec(gname, bool(g.enable_corruption), terms))
    return StepSpec(
        task=getattr(env_cfg, "task_name", type(env_cfg).__name__),
        layout=layout,
        step_dt=step_dt,
        max_episode_length=math.ceil(env_cfg.episode_length_s / step_dt),
        max_episode_length_s=env_cfg.episode_length_s,
        contact_time_abs_tol=1.0e-8,
        default_joint_pos=asset.default_joint_pos(),
        default_joint_vel=asset.default_joint_vel(),
        soft_pos_limits=asset.soft_joint_pos_limits(),
        soft_vel_limits=asset.joint_vel_limits(),
        rewards=rewards,
        dones=dones,
        obs=groups,
        command=compile_command(env_cfg.commands.base_velocity, step_dt),
        action=action,
    )

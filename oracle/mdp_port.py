"""CPU oracle: a restatement of the reference's per-step MDP pipeline in eager fp32 PyTorch.

TEST INFRASTRUCTURE - NOT PRODUCT CODE. Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; the product path (``robot_lab_b200``) never
does and fails loudly when its CUDA library is missing.

What is restated, and from where (paths relative to /root/reference; V/ =
source/robot_lab/robot_lab/tasks/manager_based/locomotion/velocity/):

  * robot_lab-owned reward terms - V/mdp/rewards.py:22-687, one function each below, citing its lines.
    PINNED: ``tests/test_oracle_vs_reference.py`` runs the *unmodified* reference file through
    ``oracle/isaaclab_shim.py`` in the build container and compares every term; the committed fixtures under
    ``tests/golden/`` were produced by the reference functions themselves (``tests/golden/make_golden.py``).
  * robot_lab-owned observation terms and command logic - V/mdp/observations.py:17-35, V/mdp/commands.py:43-85
    (the "pits" branch is identically off for the in-scope terrains, V/mdp/utils.py:27-28; it is restated
    separately in ``command_pit_restrict`` / ``is_robot_on_terrain`` for terrains that do have pits).
  * IsaacLab-owned pieces [IL] - upstream reward / observation / termination terms, the manager loops
    (RewardManager.compute, ObservationManager.compute_group, TerminationManager.compute, CommandTerm.compute,
    UniformVelocityCommand, JointAction.process_actions), math helpers and ContactSensor.compute_first_contact.
    Their source (IsaacLab v2.3.2, pinned by the reference at README.md:4,54 and
    source/robot_lab/config/extension.toml:4) is NOT vendored under /root/reference and not installed here:
    they are restated from the published upstream algorithm. **PARITY UNPINNED** for
    these pieces: the reference holds no tests, golden vectors or fixtures for them.

Arithmetic mirrors the reference's op-by-op eager evaluation (each op rounds to fp32), which also makes this
module the honest CPU baseline ("port") for bench.py.
"""

from __future__ import annotations

import math

import torch

from robot_lab_b200.spec import DoneTermSpec, ObsGroupSpec, RewardTermSpec, StepSpec

from . import philox

State = dict  # logical [N, ...] tensors, see robot_lab_b200.synthetic.make_state


# ------------------------------------------------------------------------------------------------
# isaaclab.utils.math [IL]
# ------------------------------------------------------------------------------------------------
def quat_apply(quat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    xyz = quat[:, 1:]
    t = xyz.cross(vec, dim=-1) * 2
    return vec + quat[:, 0:1] * t + xyz.cross(t, dim=-1)


def quat_apply_inverse(quat: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    xyz = quat[:, 1:]
    t = xyz.cross(vec, dim=-1) * 2
    return vec - quat[:, 0:1] * t + xyz.cross(t, dim=-1)


def yaw_quat(quat: torch.Tensor) -> torch.Tensor:
    qw, qx, qy, qz = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    yaw = torch.atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz))
    out = torch.zeros_like(quat)
    out[:, 3] = torch.sin(yaw / 2)
    out[:, 0] = torch.cos(yaw / 2)
    return out / out.norm(p=2, dim=-1, keepdim=True).clamp(min=1e-9)


def wrap_to_pi(angles: torch.Tensor) -> torch.Tensor:
    wrapped = (angles + torch.pi) % (2 * torch.pi)
    return torch.where((wrapped == 0) & (angles > 0), torch.pi, wrapped - torch.pi)


# ------------------------------------------------------------------------------------------------
# derived articulation / sensor data [IL] (ArticulationData, ContactSensor)
# ------------------------------------------------------------------------------------------------
class Derived:
    """Lazily computed quantities shared by terms (what ArticulationData properties provide upstream)."""

    def __init__(self, st: State, spec: StepSpec):
        self.st, self.spec = st, spec
        n = st["root_quat_w"].shape[0]
        self.N = n
        q = st["root_quat_w"]
        self.projected_gravity_b = quat_apply_inverse(q, torch.tensor([[0.0, 0.0, -1.0]]).repeat(n, 1))
        self.root_lin_vel_b = quat_apply_inverse(q, st["root_lin_vel_w"])
        self.root_ang_vel_b = quat_apply_inverse(q, st["root_ang_vel_w"])
        self.default_joint_pos = torch.tensor(spec.default_joint_pos, dtype=torch.float32).unsqueeze(0).repeat(n, 1)
        self.default_joint_vel = torch.tensor(spec.default_joint_vel, dtype=torch.float32).unsqueeze(0).repeat(n, 1)
        lim = torch.tensor(spec.soft_pos_limits, dtype=torch.float32)
        self.soft_joint_pos_limits = lim.unsqueeze(0).repeat(n, 1, 1)
        self.soft_joint_vel_limits = torch.tensor(spec.soft_vel_limits, dtype=torch.float32).unsqueeze(0).repeat(n, 1)
        self.terminated = st.get("terminated", torch.zeros(n, dtype=torch.bool))

    def gate(self) -> torch.Tensor:
        return torch.clamp(-self.projected_gravity_b[:, 2], 0, 0.7) / 0.7

    def first_contact(self) -> torch.Tensor:  # ContactSensor.compute_first_contact(step_dt) [IL]
        t = self.st["current_contact_time"]
        return (t > 0.0) * (t < (self.spec.step_dt + self.spec.contact_time_abs_tol))

    def first_air(self) -> torch.Tensor:
        t = self.st["current_air_time"]
        return (t > 0.0) * (t < (self.spec.step_dt + self.spec.contact_time_abs_tol))

    def heading_w(self) -> torch.Tensor:
        fwd = quat_apply(self.st["root_quat_w"], torch.tensor([[1.0, 0.0, 0.0]]).repeat(self.N, 1))
        return torch.atan2(fwd[:, 1], fwd[:, 0])


# ------------------------------------------------------------------------------------------------
# reward terms: raw value [N] (no weight, no dt)
# ------------------------------------------------------------------------------------------------
def reward_term(t: RewardTermSpec, st: State, spec: StepSpec, d: Derived | None = None) -> torch.Tensor:
    d = d or Derived(st, spec)
    fn = _REWARD_FUNCS[t.type_name]
    return fn(t, st, spec, d)


def _cmd(st):
    return st["command"]


def _rw_is_terminated(t, st, spec, d):  # [IL] rewards.is_terminated
    return d.terminated.float()


def _rw_lin_vel_z_l2(t, st, spec, d):  # V/mdp/rewards.py:647-653
    r = torch.square(d.root_lin_vel_b[:, 2])
    r *= d.gate()
    return r


def _rw_ang_vel_xy_l2(t, st, spec, d):  # :656-662
    r = torch.sum(torch.square(d.root_ang_vel_b[:, :2]), dim=1)
    r *= d.gate()
    return r


def _rw_flat_orientation_l2(t, st, spec, d):  # :678-687
    r = torch.sum(torch.square(d.projected_gravity_b[:, :2]), dim=1)
    r *= d.gate()
    return r


def _rw_base_height_l2(t, st, spec, d):  # :616-644, sensor_cfg=None branch
    r = torch.square(st["root_pos_w"][:, 2] - t.p[0])
    r *= d.gate()
    return r


def _rw_upward(t, st, spec, d):  # :608-613
    return torch.square(1 - d.projected_gravity_b[:, 2])


def _rw_joint_torques_l2(t, st, spec, d):  # [IL]
    return torch.sum(torch.square(st["applied_torque"][:, t.joint_ids]), dim=1)


def _rw_joint_vel_l2(t, st, spec, d):  # [IL]
    return torch.sum(torch.square(st["joint_vel"][:, t.joint_ids]), dim=1)


def _rw_joint_acc_l2(t, st, spec, d):  # [IL]
    return torch.sum(torch.square(st["joint_acc"][:, t.joint_ids]), dim=1)


def _joint_deviation_l1(ids, st, d):  # [IL] rewards.joint_deviation_l1
    angle = st["joint_pos"][:, ids] - d.default_joint_pos[:, ids]
    return torch.sum(torch.abs(angle), dim=1)


def _rw_joint_deviation_l1(t, st, spec, d):
    return _joint_deviation_l1(t.joint_ids, st, d)


def _rw_joint_pos_limits(t, st, spec, d):  # [IL]
    q = st["joint_pos"][:, t.joint_ids]
    out = -(q - d.soft_joint_pos_limits[:, t.joint_ids, 0]).clip(max=0.0)
    out += (q - d.soft_joint_pos_limits[:, t.joint_ids, 1]).clip(min=0.0)
    return torch.sum(out, dim=1)


def _rw_joint_vel_limits(t, st, spec, d):  # [IL]
    out = torch.abs(st["joint_vel"][:, t.joint_ids]) - d.soft_joint_vel_limits[:, t.joint_ids] * t.p[0]
    out = out.clip_(min=0.0, max=1.0)
    return torch.sum(out, dim=1)


def _rw_joint_power(t, st, spec, d):  # V/mdp/rewards.py:81-90
    return torch.sum(torch.abs(st["joint_vel"][:, t.joint_ids] * st["applied_torque"][:, t.joint_ids]), dim=1)


def _rw_stand_still(t, st, spec, d):  # :93-104
    r = _joint_deviation_l1(t.joint_ids, st, d)
    r *= torch.norm(_cmd(st), dim=1) < t.p[0]
    r *= d.gate()
    return r


def _rw_joint_pos_penalty(t, st, spec, d):  # :107-129
    cmd = torch.linalg.norm(_cmd(st), dim=1)
    body_vel = torch.linalg.norm(d.root_lin_vel_b[:, :2], dim=1)
    running = torch.linalg.norm(st["joint_pos"][:, t.joint_ids] - d.default_joint_pos[:, t.joint_ids], dim=1)
    r = torch.where(torch.logical_or(cmd > t.p[2], body_vel > t.p[1]), running, t.p[0] * running)
    r *= d.gate()
    return r


def _pairs_loop(t, values, absval):
    n = values.shape[0]
    r = torch.zeros(n)
    a, b = values[:, t.idx_a], values[:, t.idx_b]
    if absval:
        a, b = torch.abs(a), torch.abs(b)
    r += torch.sum(torch.square(a - b), dim=-1)
    return r


def _rw_joint_mirror(t, st, spec, d):  # :259-278
    r = _pairs_loop(t, st["joint_pos"], False)
    r *= t.p[0]
    r *= d.gate()
    return r


def _rw_action_mirror(t, st, spec, d):  # :281-303
    r = _pairs_loop(t, st["action"], True)
    r *= t.p[0]
    r *= d.gate()
    return r


def _rw_action_sync(t, st, spec, d):  # :306-337
    n = st["action"].shape[0]
    r = torch.zeros(n)
    for g in range(t.n_idx):
        start, size = t.idx_b[g], t.idx_c[g]
        if size < 2:
            continue
        acts = torch.abs(st["action"][:, t.idx_a[start:start + size]])
        mean = torch.mean(acts, dim=1, keepdim=True)
        r += torch.mean(torch.square(acts - mean), dim=1)
    r *= t.p[0]
    r *= d.gate()
    return r


def _rw_action_rate_l2(t, st, spec, d):  # [IL]
    return torch.sum(torch.square(st["action"] - st["prev_action"]), dim=1)


def _hist_max_norm(st, ids):
    return torch.max(torch.norm(st["net_forces_w_history"][:, :, ids], dim=-1), dim=1)[0]


def _rw_undesired_contacts(t, st, spec, d):  # :665-675
    is_contact = _hist_max_norm(st, t.body_ids) > t.p[0]
    r = torch.sum(is_contact, dim=1).float()
    r *= d.gate()
    return r


def _rw_contact_forces(t, st, spec, d):  # [IL]
    violation = _hist_max_norm(st, t.body_ids) - t.p[0]
    return torch.sum(violation.clip(min=0.0), dim=1)


def _rw_track_lin_vel_xy_exp(t, st, spec, d):  # :22-35
    err = torch.sum(torch.square(_cmd(st)[:, :2] - d.root_lin_vel_b[:, :2]), dim=1)
    r = torch.exp(-err / t.p[0])
    r *= d.gate()
    return r


def _rw_track_ang_vel_z_exp(t, st, spec, d):  # :38-48
    err = torch.square(_cmd(st)[:, 2] - d.root_ang_vel_b[:, 2])
    r = torch.exp(-err / t.p[0])
    r *= d.gate()
    return r


def _rw_track_lin_vel_xy_yaw_frame_exp(t, st, spec, d):  # :51-66
    vel_yaw = quat_apply_inverse(yaw_quat(st["root_quat_w"]), st["root_lin_vel_w"][:, :3])
    err = torch.sum(torch.square(_cmd(st)[:, :2] - vel_yaw[:, :2]), dim=1)
    r = torch.exp(-err / t.p[0])
    r *= d.gate()
    return r


def _rw_track_ang_vel_z_world_exp(t, st, spec, d):  # :69-78
    err = torch.square(_cmd(st)[:, 2] - st["root_ang_vel_w"][:, 2])
    r = torch.exp(-err / t.p[0])
    r *= d.gate()
    return r


def _rw_feet_air_time(t, st, spec, d):  # :340-360
    first = d.first_contact()[:, t.idx_a]
    r = torch.sum((st["last_air_time"][:, t.idx_a] - t.p[0]) * first, dim=1)
    r *= torch.norm(_cmd(st), dim=1) > 0.1
    r *= d.gate()
    return r


def _rw_feet_air_time_positive_biped(t, st, spec, d):  # :363-383
    air = st["current_air_time"][:, t.idx_a]
    con = st["current_contact_time"][:, t.idx_a]
    in_contact = con > 0.0
    mode = torch.where(in_contact, con, air)
    single = torch.sum(in_contact.int(), dim=1) == 1
    r = torch.min(torch.where(single.unsqueeze(-1), mode, 0.0), dim=1)[0]
    r = torch.clamp(r, max=t.p[0])
    r *= torch.norm(_cmd(st), dim=1) > 0.1
    r *= d.gate()
    return r


def _rw_feet_air_time_variance(t, st, spec, d):  # :386-397
    r = torch.var(torch.clip(st["last_air_time"][:, t.idx_a], max=0.5), dim=1) + torch.var(
        torch.clip(st["last_contact_time"][:, t.idx_a], max=0.5), dim=1)
    r *= d.gate()
    return r


def _rw_feet_gait(t, st, spec, d):  # :156-256
    air, con = st["current_air_time"], st["current_contact_time"]
    std, me2 = t.p[0], t.p[1]
    f00, f01, f10, f11 = t.idx_a[:4]

    def sync(a, b):
        se_air = torch.clip(torch.square(air[:, a] - air[:, b]), max=me2)
        se_con = torch.clip(torch.square(con[:, a] - con[:, b]), max=me2)
        return torch.exp(-(se_air + se_con) / std)

    def asyn(a, b):
        se0 = torch.clip(torch.square(air[:, a] - con[:, b]), max=me2)
        se1 = torch.clip(torch.square(con[:, a] - air[:, b]), max=me2)
        return torch.exp(-(se0 + se1) / std)

    sync_r = sync(f00, f01) * sync(f10, f11)
    async_r = asyn(f00, f10) * asyn(f01, f11) * asyn(f00, f11) * asyn(f10, f01)
    cmd = torch.linalg.norm(_cmd(st), dim=1)
    body_vel = torch.linalg.norm(d.root_lin_vel_b[:, :2], dim=1)
    r = torch.where(torch.logical_or(cmd > t.p[3], body_vel > t.p[2]), sync_r * async_r, 0.0)
    r *= d.gate()
    return r


def _rw_feet_contact(t, st, spec, d):  # :400-413
    n = torch.sum(d.first_contact()[:, t.idx_a], dim=1)
    r = (n != t.p[0]).float()
    r *= torch.linalg.norm(_cmd(st), dim=1) > 0.1
    r *= d.gate()
    return r


def _rw_feet_contact_without_cmd(t, st, spec, d):  # :416-425
    r = torch.sum(d.first_contact()[:, t.idx_a], dim=-1).float()
    r *= torch.linalg.norm(_cmd(st), dim=1) < 0.1
    r *= d.gate()
    return r


def _rw_feet_stumble(t, st, spec, d):  # :428-436 (net_forces_w = newest history sample)
    f = st["net_forces_w_history"][:, 0]
    fz = torch.abs(f[:, t.idx_c, 2])
    fxy = torch.linalg.norm(f[:, t.idx_c, :2], dim=2)
    r = torch.any(fxy > 4 * fz, dim=1).float()
    r *= d.gate()
    return r


def _feet_in_body_frame(vec_w, root_w, quat, ids):
    rel = vec_w[:, ids, :] - root_w.unsqueeze(1)
    out = torch.zeros(rel.shape[0], len(ids), 3)
    for i in range(len(ids)):
        out[:, i, :] = quat_apply_inverse(quat, rel[:, i, :])
    return out


def _rw_feet_slide(t, st, spec, d):  # :557-587
    contacts = st["net_forces_w_history"][:, :, t.idx_c, :].norm(dim=-1).max(dim=1)[0] > 1.0
    vel_b = _feet_in_body_frame(st["body_lin_vel_w"], st["root_lin_vel_w"], st["root_quat_w"], t.idx_b)
    lateral = torch.sqrt(torch.sum(torch.square(vel_b[:, :, :2]), dim=2)).view(vel_b.shape[0], -1)
    r = torch.sum(lateral * contacts, dim=1)
    r *= d.gate()
    return r


def _rw_feet_height(t, st, spec, d):  # :507-524
    err = torch.square(st["body_pos_w"][:, t.idx_b, 2] - t.p[0])
    tanh = torch.tanh(t.p[1] * torch.linalg.norm(st["body_lin_vel_w"][:, t.idx_b, :2], dim=2))
    r = torch.sum(err * tanh, dim=1)
    r *= torch.linalg.norm(_cmd(st), dim=1) > 0.1
    r *= d.gate()
    return r


def _rw_feet_height_body(t, st, spec, d):  # :527-554
    pos_b = _feet_in_body_frame(st["body_pos_w"], st["root_pos_w"], st["root_quat_w"], t.idx_b)
    vel_b = _feet_in_body_frame(st["body_lin_vel_w"], st["root_lin_vel_w"], st["root_quat_w"], t.idx_b)
    err = torch.square(pos_b[:, :, 2] - t.p[0]).view(pos_b.shape[0], -1)
    tanh = torch.tanh(t.p[1] * torch.norm(vel_b[:, :, :2], dim=2))
    r = torch.sum(err * tanh, dim=1)
    r *= torch.linalg.norm(_cmd(st), dim=1) > 0.1
    r *= d.gate()
    return r


def _rw_feet_distance_y_exp(t, st, spec, d):  # :439-461
    pos_b = _feet_in_body_frame(st["body_pos_w"], st["root_pos_w"], st["root_quat_w"], t.idx_b)
    n_feet = len(t.idx_b)
    side = torch.tensor([1.0 if i % 2 == 0 else -1.0 for i in range(n_feet)])
    desired = (t.p[0] * torch.ones(pos_b.shape[0], 1)) / 2 * side.unsqueeze(0)
    diff = torch.square(desired - pos_b[:, :, 1])
    r = torch.exp(-torch.sum(diff, dim=1) / t.p[1])
    r *= d.gate()
    return r


def _rw_feet_distance_xy_exp(t, st, spec, d):  # :464-504
    pos_b = _feet_in_body_frame(st["body_pos_w"], st["root_pos_w"], st["root_quat_w"], t.idx_b)
    n = pos_b.shape[0]
    w = t.p[0] * torch.ones(n, 1)
    l = t.p[1] * torch.ones(n, 1)
    dx = torch.cat([l / 2, l / 2, -l / 2, -l / 2], dim=1)
    dy = torch.cat([w / 2, -w / 2, w / 2, -w / 2], dim=1)
    diff = torch.square(dx - pos_b[:, :, 0]) + torch.square(dy - pos_b[:, :, 1])
    r = torch.exp(-torch.sum(diff, dim=1) / t.p[2])
    r *= d.gate()
    return r


def _rw_wheel_vel_penalty(t, st, spec, d):  # :132-153
    cmd = torch.linalg.norm(_cmd(st), dim=1)
    body_vel = torch.linalg.norm(d.root_lin_vel_b[:, :2], dim=1)
    jv = torch.abs(st["joint_vel"][:, t.idx_b])
    in_air = d.first_air()[:, t.idx_a]
    running = torch.sum(in_air * jv, dim=1)
    standing = torch.sum(jv, dim=1)
    return torch.where(torch.logical_or(cmd > t.p[1], body_vel > t.p[0]), running, standing)


_REWARD_FUNCS = {k[4:]: v for k, v in list(globals().items()) if k.startswith("_rw_")}


# ------------------------------------------------------------------------------------------------
# managers [IL]
# ------------------------------------------------------------------------------------------------
def done_term(t: DoneTermSpec, st: State, spec: StepSpec, episode_length: torch.Tensor) -> torch.Tensor:
    if t.type_name == "time_out":
        return episode_length >= spec.max_episode_length
    if t.type_name == "terrain_out_of_bounds":
        n = st["root_pos_w"].shape[0]
        if t.p[2] == 0.0:
            return torch.zeros(n, dtype=torch.bool)
        x_out = torch.abs(st["root_pos_w"][:, 0]) > t.p[0]
        y_out = torch.abs(st["root_pos_w"][:, 1]) > t.p[1]
        return torch.logical_or(x_out, y_out)
    if t.type_name == "illegal_contact":
        return torch.any(_hist_max_norm(st, t.body_ids) > t.p[0], dim=1)
    raise NotImplementedError(t.type_name)


def compute_dones(spec: StepSpec, st: State):
    """episode_length += 1; TerminationManager.compute [IL]."""
    ep = st["episode_length"] + 1
    n = ep.shape[0]
    terminated = torch.zeros(n, dtype=torch.bool)
    truncated = torch.zeros(n, dtype=torch.bool)
    bits = torch.zeros(n, dtype=torch.int32)
    for i, t in enumerate(spec.dones):
        v = done_term(t, st, spec, ep)
        if t.time_out:
            truncated |= v
        else:
            terminated |= v
        bits |= v.int() << i
    return ep, terminated, truncated, bits


def compute_rewards(spec: StepSpec, st: State, terminated: torch.Tensor):
    """RewardManager.compute(dt) [IL]: value = func * weight * dt, summed in declared order."""
    n = st["root_quat_w"].shape[0]
    st = dict(st)
    st["terminated"] = terminated
    d = Derived(st, spec)
    dt = spec.step_dt
    total = torch.zeros(n)
    sums = st["episode_sums"].clone()
    step_reward = torch.zeros(n, spec.K)
    for k, t in enumerate(spec.rewards):
        if t.weight == 0.0:
            continue
        value = reward_term(t, st, spec, d) * t.weight * dt
        total += value
        sums[:, k] += value
        step_reward[:, k] = value / dt
    return total, sums, step_reward


def command_uniforms(spec: StepSpec, st: State, rnd: dict, stream: int) -> torch.Tensor:
    """[7, N] U[0,1): given inputs (noise-as-input mode) or the kernel's Philox stream."""
    if rnd.get("cmd_uniforms") is not None:
        return rnd["cmd_uniforms"]
    n = st["root_quat_w"].shape[0]
    return philox.command_uniforms(n, rnd["seed"], rnd["step"], rnd.get("env_id_offset", 0), stream)


def _resample(spec: StepSpec, ids: torch.Tensor, u: torch.Tensor, out: dict):
    """CommandTerm._resample [IL] + UniformThresholdVelocityCommand._resample_command (V/mdp/commands.py:43-47)."""
    c = spec.command

    def uni(row, lo_hi):
        lo, hi = lo_hi
        return u[row, ids] * (hi - lo) + lo

    out["time_left"][ids] = uni(0, c.resampling_time)
    out["command"][ids, 0] = uni(1, c.lin_vel_x)
    out["command"][ids, 1] = uni(2, c.lin_vel_y)
    out["command"][ids, 2] = uni(3, c.ang_vel_z)
    if c.heading_command:
        out["heading_target"][ids] = uni(4, c.heading)
        out["is_heading_env"][ids] = u[5, ids] <= c.rel_heading_envs
    out["is_standing_env"][ids] = u[6, ids] <= c.rel_standing_envs
    out["command"][ids, :2] *= (torch.norm(out["command"][ids, :2], dim=1) > c.small_cmd_threshold).unsqueeze(1)


def compute_command(spec: StepSpec, st: State, rnd: dict, active: torch.Tensor | None = None) -> dict:
    """CommandTerm.compute(dt) [IL] for the envs in ``active`` (bool mask; None = all)."""
    c = spec.command
    n = st["root_quat_w"].shape[0]
    out = {k: st[k].clone() for k in ("command", "heading_target", "time_left", "is_heading_env", "is_standing_env",
                                      "metric_error_vel_xy", "metric_error_vel_yaw")}
    d = Derived(st, spec)
    act = torch.ones(n, dtype=torch.bool) if active is None else active
    # _update_metrics
    exy = torch.norm(st["command"][:, :2] - d.root_lin_vel_b[:, :2], dim=-1) / c.max_command_step
    eyaw = torch.abs(st["command"][:, 2] - d.root_ang_vel_b[:, 2]) / c.max_command_step
    out["metric_error_vel_xy"] = torch.where(act, st["metric_error_vel_xy"] + exy, st["metric_error_vel_xy"])
    out["metric_error_vel_yaw"] = torch.where(act, st["metric_error_vel_yaw"] + eyaw, st["metric_error_vel_yaw"])
    out["time_left"] = torch.where(act, st["time_left"] - spec.step_dt, st["time_left"])
    ids = (act & (out["time_left"] <= 0.0)).nonzero(as_tuple=False).flatten()
    if len(ids) > 0:
        _resample(spec, ids, command_uniforms(spec, st, rnd, philox.STREAM_COMMAND), out)
    # _update_command
    if c.heading_command:
        hid = (act & out["is_heading_env"]).nonzero(as_tuple=False).flatten()
        err = wrap_to_pi(out["heading_target"][hid] - d.heading_w()[hid])
        out["command"][hid, 2] = torch.clip(c.heading_control_stiffness * err, min=c.ang_vel_z[0], max=c.ang_vel_z[1])
    sid = (act & out["is_standing_env"]).nonzero(as_tuple=False).flatten()
    out["command"][sid, :] = 0.0
    return out


def obs_uniforms(spec: StepSpec, g: int, st: State, rnd: dict) -> torch.Tensor:
    key = ("obs_uniforms_policy", "obs_uniforms_critic")[g]
    if rnd.get(key) is not None:
        return rnd[key]
    n = st["root_quat_w"].shape[0]
    dims = [t.dim for t in spec.obs[g].terms]
    return philox.obs_uniforms(n, dims, g, rnd["seed"], rnd["step"], rnd.get("env_id_offset", 0))


def compute_obs_group(spec: StepSpec, g: int, st: State, rnd: dict) -> torch.Tensor:
    """ObservationManager.compute_group [IL]: clone -> +noise -> clip -> scale -> cat."""
    grp: ObsGroupSpec = spec.obs[g]
    d = Derived(st, spec)
    n = d.N
    cols = []
    u_all = None
    col0 = 0
    for t in grp.terms:
        ty = t.type_name
        if ty == "base_lin_vel":
            v = d.root_lin_vel_b
        elif ty == "base_ang_vel":
            v = d.root_ang_vel_b
        elif ty == "projected_gravity":
            v = d.projected_gravity_b
        elif ty == "generated_commands":
            v = st["command"]
        elif ty == "joint_pos_rel":
            v = st["joint_pos"][:, t.ids] - d.default_joint_pos[:, t.ids]
        elif ty == "joint_pos_rel_without_wheel":  # V/mdp/observations.py:17-27
            v = st["joint_pos"][:, t.ids] - d.default_joint_pos[:, t.ids]
            v[:, t.zero_cols] = 0
        elif ty == "joint_vel_rel":
            v = st["joint_vel"][:, t.ids] - d.default_joint_vel[:, t.ids]
        elif ty == "last_action":
            v = st["action"]
        elif ty == "height_scan":
            v = st["ray_sensor_pos_z"].unsqueeze(1) - st["ray_hits_z"] - t.p[0]
        elif ty == "phase":  # V/mdp/observations.py:30-35
            ph = st["episode_length"][:, None] * spec.step_dt / t.p[0]
            v = torch.cat([torch.sin(2 * torch.pi * ph), torch.cos(2 * torch.pi * ph)], dim=-1)
        else:
            raise NotImplementedError(ty)
        v = v.clone()
        if t.noise is not None and grp.enable_corruption:
            if u_all is None:
                u_all = obs_uniforms(spec, g, st, rnd)
            u = u_all[:, col0:col0 + t.dim]
            v = v + u * (t.noise[1] - t.noise[0]) + t.noise[0]
        if t.clip is not None:
            v = v.clip_(min=t.clip[0], max=t.clip[1])
        if t.scale is not None:
            v = v.mul_(t.scale)
        cols.append(v)
        col0 += t.dim
    return torch.cat(cols, dim=-1) if cols else torch.zeros(n, 0)


def process_action(spec: StepSpec, st: State, new_action: torch.Tensor):
    """ActionManager.process_action + JointAction.process_actions [IL]."""
    a = spec.action
    prev = st["action"].clone()
    action = new_action.clone()
    processed = action * torch.tensor(a.scale, dtype=torch.float32) + torch.tensor(a.offset, dtype=torch.float32)
    if a.clip is not None:
        clip = torch.tensor(a.clip, dtype=torch.float32)
        processed = torch.clamp(processed, min=clip[:, 0], max=clip[:, 1])
    return action, prev, processed


def step(spec: StepSpec, st: State, rnd: dict, skip_done_envs: bool = False) -> dict:
    """ManagerBasedRLEnv.step() [IL] minus physics and the external reset.

    With ``skip_done_envs`` the command update is withheld from envs that are done this step, exactly like the
    fused kernel's RL_PHASE_SKIP_DONE_ENVS (they are refreshed after the reset by :func:`refresh_after_reset`).
    """
    ep, terminated, truncated, bits = compute_dones(spec, st)
    total, sums, step_reward = compute_rewards(spec, st, terminated)
    done = terminated | truncated
    reset_ids = done.nonzero(as_tuple=False).flatten().to(torch.int32)
    cmd = compute_command(spec, st, rnd, active=(~done) if skip_done_envs else None)
    st2 = dict(st)
    st2.update(cmd)
    st2["episode_length"] = ep
    return {
        "episode_length": ep, "terminated": terminated, "truncated": truncated, "done_bits": bits,
        "reward": total, "episode_sums": sums, "step_reward": step_reward, "reset_ids": reset_ids,
        **cmd,
        "obs_policy": compute_obs_group(spec, 0, st2, rnd),
        "obs_critic": compute_obs_group(spec, 1, st2, rnd),
    }


def quat_from_euler_xyz(roll: torch.Tensor, pitch: torch.Tensor, yaw: torch.Tensor) -> torch.Tensor:
    """isaaclab.utils.math.quat_from_euler_xyz [IL] -> (w, x, y, z)."""
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qw, qx, qy, qz], dim=-1)


def quat_mul(q1: torch.Tensor, q2: torch.Tensor) -> torch.Tensor:
    """isaaclab.utils.math.quat_mul [IL]."""
    w1, x1, y1, z1 = q1[..., 0], q1[..., 1], q1[..., 2], q1[..., 3]
    w2, x2, y2, z2 = q2[..., 0], q2[..., 1], q2[..., 2], q2[..., 3]
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return torch.stack([w, x, y, z], dim=-1)


def reset_scene_state(spec: StepSpec, st: State, ids: torch.Tensor, cfg, env_origins: torch.Tensor,
                      uniforms: torch.Tensor, assigned_to_pits: torch.Tensor | None = None) -> dict:
    """``reset_root_state_uniform`` (V/mdp/events.py:205-271, non-pit branch) followed by ``reset_joints_by_scale``
    [IL] (wired at V/velocity_env_cfg.py:326-363) for ``ids``. ``uniforms`` is [12 + 2J, N] U[0,1): pose 6, velocity 6,
    joint position J, joint velocity J; ``sample_uniform`` [IL] = rand * (hi - lo) + lo. ``cfg``: cfg.ResetStateCfg."""
    ids = ids.long()
    J = spec.J
    out = {k: st[k].clone() for k in ("root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w", "joint_pos", "joint_vel")}
    keys = ("x", "y", "z", "roll", "pitch", "yaw")
    drs = torch.tensor([0.0, 0.0, spec.layout.asset.init_root_height, 1.0, 0.0, 0.0, 0.0] + [0.0] * 6)
    root = drs.unsqueeze(0).repeat(len(ids), 1)
    pr = torch.tensor([cfg.pose_range.get(k, (0.0, 0.0)) for k in keys])
    vr = torch.tensor([cfg.velocity_range.get(k, (0.0, 0.0)) for k in keys])
    up, uv = uniforms[0:6, ids].t(), uniforms[6:12, ids].t()
    pose = up * (pr[:, 1] - pr[:, 0]) + pr[:, 0]
    vel = uv * (vr[:, 1] - vr[:, 0]) + vr[:, 0]
    out["root_pos_w"][ids] = root[:, 0:3] + env_origins[ids] + pose[:, 0:3]
    out["root_quat_w"][ids] = quat_mul(root[:, 3:7], quat_from_euler_xyz(pose[:, 3], pose[:, 4], pose[:, 5]))
    velocities = root[:, 7:13] + vel
    out["root_lin_vel_w"][ids] = velocities[:, 0:3]
    out["root_ang_vel_w"][ids] = velocities[:, 3:6]
    q0, qd0 = torch.tensor(spec.default_joint_pos), torch.tensor(spec.default_joint_vel)
    lim = torch.tensor(spec.soft_pos_limits)
    vlim = torch.tensor(spec.soft_vel_limits)
    plo, phi = cfg.joint_position_range
    vlo, vhi = cfg.joint_velocity_range
    jp = q0.unsqueeze(0) * (uniforms[12:12 + J, ids].t() * (phi - plo) + plo)
    jv = qd0.unsqueeze(0) * (uniforms[12 + J:12 + 2 * J, ids].t() * (vhi - vlo) + vlo)
    if assigned_to_pits is not None:
        # V/mdp/events.py:232-244: envs assigned to the "pits" sub-terrain get the default root state at their origin
        pit = ids[assigned_to_pits[ids].bool()]
        if len(pit) > 0:
            rp = drs.unsqueeze(0).repeat(len(pit), 1)
            out["root_pos_w"][pit] = rp[:, 0:3] + env_origins[pit]
            out["root_quat_w"][pit] = rp[:, 3:7]
            out["root_lin_vel_w"][pit] = 0.0
            out["root_ang_vel_w"][pit] = 0.0
    out["joint_pos"][ids] = torch.maximum(torch.minimum(jp, lim[:, 1]), lim[:, 0])
    out["joint_vel"][ids] = torch.maximum(torch.minimum(jv, vlim), -vlim)
    return out


def contact_sensor_update(spec: StepSpec, st: State, net_forces_w: torch.Tensor, dt: float,
                          force_threshold: float = 1.0) -> dict:
    """ContactSensor._update_buffers_impl [IL] (isaaclab/sensors/contact_sensor/contact_sensor.py, IsaacLab v2.3.2;
    configured at V/velocity_env_cfg.py:86 with history_length=3, track_air_time=True, updated every physics
    sub-step, :726). PARITY UNPINNED: the file is not vendored; restated from the upstream algorithm.

    ``net_forces_w`` is [N, B, 3] in the history body space; the timers live in the (feet) timer body space."""
    hist = st["net_forces_w_history"]
    out = {}
    new_hist = torch.roll(hist, 1, dims=1)
    new_hist[:, 0] = net_forces_w
    out["net_forces_w_history"] = new_hist
    names_h = list(spec.layout.hist_body_names)
    t2h = torch.tensor([names_h.index(n) for n in spec.layout.time_body_names], dtype=torch.long)
    is_contact = torch.norm(net_forces_w[:, t2h, :], dim=-1) > force_threshold
    cur_air, cur_con = st["current_air_time"], st["current_contact_time"]
    is_first_contact = (cur_air > 0) * is_contact
    is_first_detached = (cur_con > 0) * ~is_contact
    out["last_air_time"] = torch.where(is_first_contact, cur_air + dt, st["last_air_time"])
    out["current_air_time"] = torch.where(~is_contact, cur_air + dt, torch.zeros(()))
    out["last_contact_time"] = torch.where(is_first_detached, cur_con + dt, st["last_contact_time"])
    out["current_contact_time"] = torch.where(is_contact, cur_con + dt, torch.zeros(()))
    return out


def reset_envs(spec: StepSpec, st: State, ids: torch.Tensor, done_bits: torch.Tensor, rnd: dict) -> tuple[dict, dict]:
    """Manager part of ManagerBasedRLEnv._reset_idx [IL]: logging means, zeroing, command resample."""
    ids = ids.long()
    out = {k: st[k].clone() for k in ("episode_sums", "action", "prev_action", "command", "heading_target", "time_left",
                                      "is_heading_env", "is_standing_env", "metric_error_vel_xy",
                                      "metric_error_vel_yaw", "episode_length")}
    log = {
        "episode_sum_mean": torch.stack([torch.mean(st["episode_sums"][ids, k]) for k in range(spec.K)])
        if len(ids) > 0 else torch.zeros(spec.K),
        "done_term_count": torch.tensor([float(torch.count_nonzero((done_bits[ids] >> i) & 1)) for i in range(8)]),
        "metric_mean": torch.stack([torch.mean(st["metric_error_vel_xy"][ids]), torch.mean(st["metric_error_vel_yaw"][ids])])
        if len(ids) > 0 else torch.zeros(2),
    }
    out["episode_sums"][ids] = 0.0
    out["action"][ids] = 0.0
    out["prev_action"][ids] = 0.0
    out["metric_error_vel_xy"][ids] = 0.0
    out["metric_error_vel_yaw"][ids] = 0.0
    out["episode_length"][ids] = 0
    if len(ids) > 0:
        _resample(spec, ids, command_uniforms(spec, st, rnd, philox.STREAM_RESET_COMMAND), out)
    return out, log


# ------------------------------------------------------------------------------------------------
# actuator models between process_action and physics
# ------------------------------------------------------------------------------------------------
def actuator_step(table: dict, joint_pos_target: torch.Tensor, joint_pos: torch.Tensor, joint_vel: torch.Tensor,
                  joint_vel_target: torch.Tensor | None = None, joint_effort_target: torch.Tensor | None = None):
    """Articulation._apply_actuator_model [IL] for every actuator group of the asset (isaaclab/assets/articulation/
    articulation.py; isaaclab/actuators/actuator_pd.py: ImplicitActuator / IdealPDActuator / DCMotor, IsaacLab
    v2.3.2). The reference selects the models at assets/unitree.py:55-63 (A1, DCMotor), :107-115 (Go2, DCMotor),
    :504-621 (G1, ImplicitActuator). PARITY UNPINNED: the file is not vendored; restated from the upstream algorithm
    (the v2.2+ four-quadrant form of DCMotor._clip_effort).

    ``table`` = RobotAsset.actuator_table(). Returns (computed_torque, applied_torque), [N, J]; joints without an
    actuator keep 0."""
    n, J = joint_pos.shape
    f32 = torch.float32
    kp = torch.tensor(table["stiffness"], dtype=f32)
    kd = torch.tensor(table["damping"], dtype=f32)
    lim = torch.tensor(table["effort_limit"], dtype=f32).unsqueeze(0).repeat(n, 1)   # parsed into [N, J] tensors [IL]
    vlim = torch.tensor(table["velocity_limit"], dtype=f32).unsqueeze(0).repeat(n, 1)
    vt = torch.zeros_like(joint_pos) if joint_vel_target is None else joint_vel_target
    et = torch.zeros_like(joint_pos) if joint_effort_target is None else joint_effort_target
    # IdealPDActuator.compute / ImplicitActuator.compute
    error_pos = joint_pos_target - joint_pos
    error_vel = vt - joint_vel
    computed = kp * error_pos + kd * error_vel + et
    applied = torch.clip(computed, min=-lim, max=lim)
    # DCMotor._clip_effort
    for j, kind in enumerate(table["kind"]):
        if kind == "dc_motor":
            sat = float(table["saturation_effort"][j])
            vel_at_effort_lim = vlim[:, j] * (1 + lim[:, j] / sat)
            vel = torch.clip(joint_vel[:, j], min=-vel_at_effort_lim, max=vel_at_effort_lim)
            torque_speed_top = sat * (1.0 - vel / vlim[:, j])
            torque_speed_bottom = sat * (-1.0 - vel / vlim[:, j])
            max_effort = torch.clip(torque_speed_top, max=lim[:, j])
            min_effort = torch.clip(torque_speed_bottom, min=-lim[:, j])
            applied[:, j] = torch.clip(computed[:, j], min=min_effort, max=max_effort)
        elif kind == "none":
            computed[:, j] = 0.0
            applied[:, j] = 0.0
    return computed, applied


# ------------------------------------------------------------------------------------------------
# terrain-aware command restriction (V/mdp/utils.py, V/mdp/commands.py:49-85)
# ------------------------------------------------------------------------------------------------
def terrain_column_range(sub_terrains: list[str], proportions: list[float], num_cols: int, name: str):
    """``_get_terrain_column_range`` (V/mdp/utils.py:16-41): fp32 torch cumsum of the normalised proportions, python
    ``round`` of (cumsum * num_cols)."""
    if name not in sub_terrains:
        return None
    p = torch.tensor(proportions)
    p = p / p.sum()
    c = torch.cumsum(p, dim=0)
    i = sub_terrains.index(name)
    col_start = round((0.0 if i == 0 else c[i - 1].item()) * num_cols)
    col_end = round(c[i].item() * num_cols)
    return col_start, col_end


def is_robot_on_terrain(root_pos_w: torch.Tensor, terrain_origins: torch.Tensor, col_range) -> torch.Tensor:
    """``is_robot_on_terrain`` (V/mdp/utils.py:73-127) for a generator terrain: argmin over the xy distance to every
    terrain origin (flat row-major index, first minimum), column = index % num_cols, inside [col_start, col_end).
    The distance is evaluated as sqrt(dx^2 + dy^2) - the definition of torch.cdist; the reference call goes through
    cdist's matmul path for these sizes, which rounds differently (robots within ~1e-4 m of a cell boundary)."""
    n = root_pos_w.shape[0]
    if col_range is None:
        return torch.zeros(n, dtype=torch.bool)
    rows, cols, _ = terrain_origins.shape
    o = terrain_origins[:, :, :2].reshape(rows * cols, 2)
    dx = root_pos_w[:, 0:1] - o[:, 0].unsqueeze(0)
    dy = root_pos_w[:, 1:2] - o[:, 1].unsqueeze(0)
    d = torch.sqrt(dx * dx + dy * dy)
    col = torch.argmin(d, dim=1) % cols
    return (col >= col_range[0]) & (col < col_range[1])


def command_pit_restrict(spec: StepSpec, st: State, on_pits: torch.Tensor, was_on_pit: torch.Tensor,
                         u: torch.Tensor) -> dict:
    """The tail of ``UniformThresholdVelocityCommand._update_command`` (V/mdp/commands.py:61-85) applied to the
    command state in ``st`` (i.e. after the parent's _update_command). ``u`` is the [7, N] uniform table (row 0
    unused: _resample_command does not touch time_left)."""
    c = spec.command
    out = {k: st[k].clone() for k in ("command", "heading_target", "is_heading_env", "is_standing_env")}
    left = (was_on_pit & ~on_pits).nonzero(as_tuple=False).flatten()
    if len(left) > 0:
        tmp = dict(out)
        tmp["time_left"] = torch.zeros(on_pits.shape[0])
        _resample(spec, left, u, tmp)
    pit = on_pits.nonzero(as_tuple=False).flatten()
    if len(pit) > 0:
        out["command"][pit, 0] = torch.clamp(torch.abs(out["command"][pit, 0]), min=0.3, max=0.6)
        out["command"][pit, 1] = 0.0
        out["command"][pit, 2] = 0.0
        if c.heading_command:
            out["heading_target"][pit] = 0.0
    out["was_on_pit"] = on_pits.clone()
    return out


def grid_pattern_ray_starts(size: tuple[float, float], resolution: float, offset_z: float) -> torch.Tensor:
    """patterns.grid_pattern [IL] with ordering "xy" + RayCaster._initialize_rays_impl [IL]: [R, 3] ray starts in the
    sensor frame, x fastest, the sensor offset (0, 0, offset_z) already added (V/velocity_env_cfg.py:70-77)."""
    x = torch.arange(start=-size[0] / 2, end=size[0] / 2 + 1.0e-9, step=resolution)
    y = torch.arange(start=-size[1] / 2, end=size[1] / 2 + 1.0e-9, step=resolution)
    gx, gy = torch.meshgrid(x, y, indexing="xy")
    starts = torch.zeros(gx.numel(), 3)
    starts[:, 0] = gx.flatten()
    starts[:, 1] = gy.flatten()
    starts[:, 2] += offset_z
    return starts


def height_scan_cast(heights: torch.Tensor, x0: float, y0: float, horizontal_scale: float, ray_starts: torch.Tensor,
                     root_pos_w: torch.Tensor, root_quat_w: torch.Tensor):
    """RayCaster._update_buffers_impl [IL] with ray_alignment="yaw" over the triangle mesh of a height field
    (isaaclab/terrains/height_field/utils.py: convert_height_field_to_mesh - vertex (i, j) at (i, j) * scale, two
    triangles per cell, (v00, v11, v01) and (v00, v10, v11)). A vertical ray hits the mesh at the piecewise-linear
    interpolation of the vertex heights; rays outside the field hit nothing (+inf). PARITY UNPINNED (upstream casts
    with Warp's mesh query; equal up to the rounding of t * direction, ~1e-5 m for 20 m rays).

    Returns (ray_hits_z [N, R], sensor_pos_z [N])."""
    n, R = root_pos_w.shape[0], ray_starts.shape[0]
    q = yaw_quat(root_quat_w)
    starts = quat_apply(q.repeat_interleave(R, dim=0), ray_starts.repeat(n, 1)).reshape(n, R, 3)
    wx = starts[:, :, 0] + root_pos_w[:, 0:1]
    wy = starts[:, :, 1] + root_pos_w[:, 1:2]
    gx = (wx - x0) / horizontal_scale
    gy = (wy - y0) / horizontal_scale
    nx, ny = heights.shape
    inside = (gx >= 0) & (gy >= 0) & (gx <= nx - 1) & (gy <= ny - 1)
    ix = gx.clamp(0, nx - 1).long().clamp(max=nx - 2)
    iy = gy.clamp(0, ny - 1).long().clamp(max=ny - 2)
    fx = gx - ix.float()
    fy = gy - iy.float()
    h00, h01 = heights[ix, iy], heights[ix, iy + 1]
    h10, h11 = heights[ix + 1, iy], heights[ix + 1, iy + 1]
    za = (h00 + fx * (h11 - h01)) + fy * (h01 - h00)
    zb = (h00 + fx * (h10 - h00)) + fy * (h11 - h10)
    z = torch.where(fy >= fx, za, zb)
    z = torch.where(inside, z, torch.full_like(z, float("inf")))
    return z, root_pos_w[:, 2].clone()

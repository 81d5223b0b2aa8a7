"""Seeded synthetic env state for benchmarks and parity tests (SURVEY.md 8(d) "Synthetic inputs", Appendix C).

There is no physics engine in this tier: the state the MDP step reads (root pose / velocities, joint state,
contact history, air-time timers, feet kinematics, height-scan hits) and the manager state it updates are
drawn from distributions chosen so that every threshold on the path is crossed by a non-trivial fraction of envs
(contact 1 N / 100 N, |cmd| 0.1 / 0.2, |v_xy| 0.5, soft joint limits, time-out at max_episode_length, terrain
bounds, +inf ray hits, first-contact windows of exactly step_dt).

Tensors are produced in their *logical* IsaacLab shapes ([N, ...], fp32, CPU); ``robot_lab_b200.state`` lays them
out on the device (SoA for small per-env fields, AoS rows for the sensor streams).
"""

from __future__ import annotations

import math

import torch

from .spec import StepSpec


def _quat_from_axis_angle(axis: torch.Tensor, angle: torch.Tensor) -> torch.Tensor:
    axis = axis / axis.norm(dim=-1, keepdim=True).clamp(min=1e-9)
    half = 0.5 * angle
    return torch.cat([torch.cos(half).unsqueeze(-1), axis * torch.sin(half).unsqueeze(-1)], dim=-1)


def make_state(spec: StepSpec, num_envs: int, seed: int = 1234, rank: int = 0) -> dict[str, torch.Tensor]:
    """Physics / sensor state + manager state + random inputs for one step, logical shapes, CPU fp32."""
    g = torch.Generator().manual_seed(seed + rank)
    N, J, B, T, Bt, Ba, R, A, K = num_envs, spec.J, spec.B, spec.T, spec.Bt, spec.Ba, spec.R, spec.A, spec.K
    f32 = torch.float32

    def U(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=f32) * (hi - lo) + lo

    def Nrm(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=f32) * std

    def frac(p):
        return torch.rand(N, generator=g) < p

    s: dict[str, torch.Tensor] = {}
    # ---- root ----
    quat = Nrm(N, 4)
    quat = quat / quat.norm(dim=-1, keepdim=True).clamp(min=1e-9)
    upright = frac(0.55)  # most robots walk roughly upright: yaw + small tilt
    yaw = _quat_from_axis_angle(torch.tensor([[0.0, 0.0, 1.0]]).expand(N, 3), U(N, lo=-math.pi, hi=math.pi))
    tilt = _quat_from_axis_angle(Nrm(N, 3) * torch.tensor([1.0, 1.0, 0.0]) + 1e-6, Nrm(N, std=0.25))
    yq = _quat_mul(yaw, tilt)
    quat = torch.where(upright.unsqueeze(-1), yq, quat)
    inverted = frac(0.02)
    flip = torch.tensor([[0.0, 1.0, 0.0, 0.0]]).expand(N, 4)  # 180 deg about x: gravity gate = 0
    quat = torch.where(inverted.unsqueeze(-1), _quat_mul(yaw, flip), quat)
    s["root_quat_w"] = quat.contiguous()
    pos = torch.stack([U(N, lo=-60.0, hi=60.0), U(N, lo=-100.0, hi=100.0), U(N, lo=0.2, hi=0.6)], dim=-1)
    s["root_pos_w"] = pos
    s["root_lin_vel_w"] = Nrm(N, 3)
    s["root_ang_vel_w"] = Nrm(N, 3)
    slow = frac(0.25)  # |v_xy| < 0.5 branch of joint_pos_penalty / feet_gait
    s["root_lin_vel_w"] = torch.where(slow.unsqueeze(-1), s["root_lin_vel_w"] * 0.2, s["root_lin_vel_w"])
    # ---- joints ----
    dj = torch.tensor(spec.default_joint_pos, dtype=f32)
    lim = torch.tensor(spec.soft_pos_limits, dtype=f32)  # [J, 2]
    jp = dj + U(N, J, lo=-0.5, hi=0.5)
    over = torch.rand(N, J, generator=g) < 0.03
    side = torch.rand(N, J, generator=g) < 0.5
    beyond = torch.where(side, lim[:, 0] - U(N, J, lo=0.0, hi=0.2), lim[:, 1] + U(N, J, lo=0.0, hi=0.2))
    s["joint_pos"] = torch.where(over, beyond, jp)
    s["joint_vel"] = Nrm(N, J, std=3.0)
    s["joint_acc"] = Nrm(N, J, std=50.0)
    s["applied_torque"] = U(N, J, lo=-23.5, hi=23.5)
    # ---- contact-force history [N, T, B, 3] ----
    mag = -30.0 * torch.log1p(-U(N, T, B))  # Exp(mean 30 N)
    mag = torch.where(torch.rand(N, T, B, generator=g) < 0.3, mag, torch.zeros(()))
    dirn = Nrm(N, T, B, 3)
    dirn = dirn / dirn.norm(dim=-1, keepdim=True).clamp(min=1e-9)
    hist = dirn * mag.unsqueeze(-1)
    if N >= 8 and B > 0:  # exactly-on-threshold samples: |F| == 1.0 is NOT a contact (strict >)
        hist[0] = 0.0
        hist[0, 0, :, 0] = 1.0
        hist[1] = 0.0
        hist[1, 1, :, 2] = -100.0
    s["net_forces_w_history"] = hist.contiguous()
    # ---- air / contact timers (mutually exclusive per foot) ----
    in_contact = torch.rand(N, Bt, generator=g) < 0.5
    tcur = U(N, Bt)
    dt32 = torch.tensor(spec.step_dt, dtype=f32)
    tcur = torch.where(torch.rand(N, Bt, generator=g) < 0.05, dt32, tcur)  # first-contact / first-air window
    s["current_contact_time"] = torch.where(in_contact, tcur, torch.zeros(()))
    s["current_air_time"] = torch.where(in_contact, torch.zeros(()), tcur)
    s["last_air_time"] = U(N, Bt)
    s["last_contact_time"] = U(N, Bt)
    # ---- feet kinematics ----
    s["body_pos_w"] = pos.unsqueeze(1) + Nrm(N, Ba, 3, std=0.3) - torch.tensor([0.0, 0.0, 0.25])
    s["body_lin_vel_w"] = s["root_lin_vel_w"].unsqueeze(1) + Nrm(N, Ba, 3)
    # ---- height scan ----
    sensor_z = pos[:, 2] + 20.0
    s["ray_sensor_pos_z"] = sensor_z.contiguous()
    if R > 0:
        hits = sensor_z.unsqueeze(1) - 0.5 - U(N, R, lo=-1.2, hi=1.2)
        hits = torch.where(torch.rand(N, R, generator=g) < 0.005, torch.full((), float("inf")), hits)
        s["ray_hits_z"] = hits.contiguous()
    else:
        s["ray_hits_z"] = torch.zeros(N, 0, dtype=f32)
    # ---- manager state ----
    s["action"] = Nrm(N, A).clamp(-100.0, 100.0)
    s["prev_action"] = Nrm(N, A).clamp(-100.0, 100.0)
    s["new_action"] = Nrm(N, A).clamp(-100.0, 100.0)
    cmd = U(N, 3, lo=-1.0, hi=1.0)
    keep = (cmd[:, :2].norm(dim=1) > spec.command.small_cmd_threshold).unsqueeze(1)
    cmd[:, :2] = cmd[:, :2] * keep
    tiny = frac(0.10)  # |cmd| < 0.1 branch (stand_still, feet_contact_without_cmd)
    cmd = torch.where(tiny.unsqueeze(-1), cmd * 0.02, cmd)
    s["command"] = cmd.contiguous()
    s["heading_target"] = U(N, lo=-math.pi, hi=math.pi)
    tl = U(N, lo=0.0, hi=10.0)
    s["time_left"] = torch.where(frac(0.03), torch.full((), 0.01), tl)  # resample fires
    s["is_heading_env"] = frac(0.9)
    s["is_standing_env"] = frac(0.02)
    s["metric_error_vel_xy"] = U(N)
    s["metric_error_vel_yaw"] = U(N)
    ep = torch.randint(0, spec.max_episode_length, (N,), generator=g, dtype=torch.int32)
    ep = torch.where(frac(0.015), torch.full((), spec.max_episode_length - 1, dtype=torch.int32), ep)
    s["episode_length"] = ep
    s["episode_sums"] = Nrm(N, K, std=0.1)
    # ---- random inputs (noise-as-input mode) ----
    s["cmd_uniforms"] = U(7, N)
    s["obs_uniforms_policy"] = U(N, max(spec.obs[0].dim, 1))[:, : spec.obs[0].dim].contiguous()
    s["obs_uniforms_critic"] = U(N, max(spec.obs[1].dim, 1))[:, : spec.obs[1].dim].contiguous()
    return s


def _quat_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    w1, x1, y1, z1 = a.unbind(-1)
    w2, x2, y2, z2 = b.unbind(-1)
    return torch.stack([
        w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
        w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
        w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
    ], dim=-1)


STATE_KEYS = (
    "root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w", "joint_pos", "joint_vel", "joint_acc",
    "applied_torque", "net_forces_w_history", "current_air_time", "last_air_time", "current_contact_time",
    "last_contact_time", "body_pos_w", "body_lin_vel_w", "ray_hits_z", "ray_sensor_pos_z",
)
MDP_KEYS = (
    "action", "prev_action", "command", "heading_target", "time_left", "is_heading_env", "is_standing_env",
    "metric_error_vel_xy", "metric_error_vel_yaw", "episode_length", "episode_sums",
)

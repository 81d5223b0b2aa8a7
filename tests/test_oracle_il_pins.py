"""Independent answers for the IsaacLab-owned arithmetic the oracle restates (isaaclab.utils.math, ArticulationData,
DCMotor - not vendored by the reference): every helper below is pinned against a source that
does NOT share code or authorship with oracle/mdp_port.py - scipy.spatial.transform.Rotation (float64), closed forms,
and hand-worked torque-speed cases. What stays unpinned after this file: the manager loops and the upstream term
formulas themselves (COVERAGE.md).

Quaternion convention: IsaacLab is (w, x, y, z), scipy is (x, y, z, w).
"""

import math

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

from oracle import mdp_port as port

N = 4096


def _rand_quats(seed=0, n=N):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    # corner cases: identity, the three half turns, near-upside-down, tiny rotations
    q[0] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    q[1] = torch.tensor([0.0, 1.0, 0.0, 0.0])
    q[2] = torch.tensor([0.0, 0.0, 1.0, 0.0])
    q[3] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    q[4] = torch.tensor([1e-4, 1.0, 0.0, 0.0]) / math.sqrt(1 + 1e-8)
    q[5] = torch.tensor([1.0, 1e-4, -1e-4, 1e-4]) / math.sqrt(1 + 3e-8)
    q[6] = -q[7]                                        # q and -q are the same rotation
    return q


def _scipy(q: torch.Tensor) -> Rotation:
    w, x, y, z = (q[:, i].double().numpy() for i in range(4))
    return Rotation.from_quat(np.stack([x, y, z, w], axis=-1))


def test_quat_apply_and_inverse_against_scipy():
    q = _rand_quats(1)
    v = torch.randn(N, 3, generator=torch.Generator().manual_seed(2)) * 3.0
    R = _scipy(q)
    want = torch.from_numpy(R.apply(v.double().numpy())).float()
    want_inv = torch.from_numpy(R.inv().apply(v.double().numpy())).float()
    torch.testing.assert_close(port.quat_apply(q, v), want, rtol=1e-5, atol=5e-6)
    torch.testing.assert_close(port.quat_apply_inverse(q, v), want_inv, rtol=1e-5, atol=5e-6)
    # the two are inverses of each other
    torch.testing.assert_close(port.quat_apply_inverse(q, port.quat_apply(q, v)), v, rtol=1e-5, atol=1e-5)


def test_projected_gravity_and_base_frame_velocities_against_scipy():
    """ArticulationData [IL]: projected_gravity_b = R^T (0, 0, -1), root_lin_vel_b = R^T v_w, root_ang_vel_b = R^T w_w."""
    q = _rand_quats(3)
    R = _scipy(q)
    g = torch.tensor([[0.0, 0.0, -1.0]]).repeat(N, 1)
    want = torch.from_numpy(R.inv().apply(g.double().numpy())).float()
    torch.testing.assert_close(port.quat_apply_inverse(q, g), want, rtol=1e-5, atol=2e-6)
    assert port.quat_apply_inverse(q[:1], g[:1])[0].tolist() == [0.0, 0.0, -1.0]        # identity: exactly (0, 0, -1)
    up = port.quat_apply_inverse(q[1:2], g[1:2])[0]                                      # half turn about x: gravity points up
    assert up[2].item() == pytest.approx(1.0, abs=1e-6)


def test_yaw_quat_against_scipy_euler():
    """yaw_quat keeps only the rotation about world z: compare with scipy's intrinsic ZYX yaw."""
    q = _rand_quats(4)
    got = port.yaw_quat(q)
    yaw = _scipy(q).as_euler("ZYX")[:, 0]            # yaw, pitch, roll
    want = np.stack([np.cos(yaw / 2), np.zeros_like(yaw), np.zeros_like(yaw), np.sin(yaw / 2)], axis=-1)
    # away from the gimbal singularity (|pitch| -> 90 deg) where yaw itself is ill-conditioned
    pitch = _scipy(q).as_euler("ZYX")[:, 1]
    ok = torch.from_numpy(np.abs(np.abs(pitch) - np.pi / 2) > 1e-2)
    g, w = got[ok].double(), torch.from_numpy(want)[ok]
    sign = torch.sign((g * w).sum(-1, keepdim=True))     # q and -q are the same rotation
    torch.testing.assert_close(g * sign, w, rtol=0, atol=2e-5)
    assert torch.all(got[:, 1] == 0) and torch.all(got[:, 2] == 0)
    torch.testing.assert_close(got.norm(dim=-1), torch.ones(N), rtol=0, atol=1e-6)


def test_heading_w_against_scipy():
    """ArticulationData.heading_w [IL] = atan2 of the rotated x axis."""
    q = _rand_quats(5)
    st = {"root_quat_w": q}
    d = port.Derived.__new__(port.Derived)
    d.st, d.N = st, N
    fwd = _scipy(q).apply(np.array([1.0, 0.0, 0.0]))
    want = np.arctan2(fwd[:, 1], fwd[:, 0])
    got = port.Derived.heading_w(d).double().numpy()
    flat = np.hypot(fwd[:, 0], fwd[:, 1]) > 1e-3          # heading of a vertical x axis is undefined
    diff = np.abs(np.angle(np.exp(1j * (got - want))))
    assert diff[flat].max() < 2e-4


def test_quat_from_euler_xyz_against_scipy():
    g = torch.Generator().manual_seed(6)
    r, p, y = ((torch.rand(N, generator=g) * 2 - 1) * math.pi for _ in range(3))
    r[0], p[0], y[0] = 0.0, 0.0, 0.0
    r[1], p[1], y[1] = 0.3, 0.0, 0.0
    got = port.quat_from_euler_xyz(r, p, y).double()
    # IsaacLab: extrinsic x-y-z (roll about x, then pitch about y, then yaw about z) = scipy "xyz" (lower case)
    R = Rotation.from_euler("xyz", np.stack([r.double().numpy(), p.double().numpy(), y.double().numpy()], axis=-1))
    xyzw = R.as_quat()
    want = torch.from_numpy(np.concatenate([xyzw[:, 3:4], xyzw[:, :3]], axis=-1))
    sign = torch.sign((got * want).sum(-1, keepdim=True))
    torch.testing.assert_close(got * sign, want, rtol=0, atol=3e-6)
    assert got[0].tolist() == [1.0, 0.0, 0.0, 0.0]


def test_quat_mul_against_scipy():
    q1, q2 = _rand_quats(7), _rand_quats(8)
    got = port.quat_mul(q1, q2).double()
    xyzw = (_scipy(q1) * _scipy(q2)).as_quat()       # composition: apply q2 first, then q1 = Hamilton product q1 * q2
    want = torch.from_numpy(np.concatenate([xyzw[:, 3:4], xyzw[:, :3]], axis=-1))
    sign = torch.sign((got * want).sum(-1, keepdim=True))
    torch.testing.assert_close(got * sign, want, rtol=0, atol=3e-6)
    # textbook Hamilton product, written out independently in float64
    a, b = q1.double(), q2.double()
    w = a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2] - a[:, 3] * b[:, 3]
    x = a[:, 0] * b[:, 1] + a[:, 1] * b[:, 0] + a[:, 2] * b[:, 3] - a[:, 3] * b[:, 2]
    yv = a[:, 0] * b[:, 2] - a[:, 1] * b[:, 3] + a[:, 2] * b[:, 0] + a[:, 3] * b[:, 1]
    z = a[:, 0] * b[:, 3] + a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1] + a[:, 3] * b[:, 0]
    torch.testing.assert_close(got, torch.stack([w, x, yv, z], dim=-1), rtol=0, atol=3e-6)


def test_wrap_to_pi_closed_form():
    """isaaclab.utils.math.wrap_to_pi: into (-pi, pi], with +pi (not -pi) for positive odd multiples of pi."""
    pi = math.pi
    f32pi = float(torch.tensor(pi, dtype=torch.float32))
    cases = {0.0: 0.0, 1.0: 1.0, -1.0: -1.0, 3.0: 3.0, -3.0: -3.0, 4.0: 4.0 - 2 * pi, -4.0: -4.0 + 2 * pi,
             7.0: 7.0 - 2 * pi, -7.0: -7.0 + 2 * pi, 100.0: 100.0 - 32 * pi, -100.0: -100.0 + 32 * pi}
    a = torch.tensor(list(cases.keys()), dtype=torch.float32)
    torch.testing.assert_close(port.wrap_to_pi(a), torch.tensor(list(cases.values()), dtype=torch.float32), rtol=0, atol=2e-5)
    # exactly +-pi and multiples of 2 pi (as float32 values)
    edge = torch.tensor([f32pi, -f32pi, 2 * f32pi, -2 * f32pi, 3 * f32pi, 0.0], dtype=torch.float32)
    got = port.wrap_to_pi(edge)
    assert got[0].item() == pytest.approx(pi, abs=1e-6)            # +pi stays +pi
    assert abs(got[1].item()) == pytest.approx(pi, abs=1e-6)       # -pi is the same angle
    assert abs(got[2].item()) < 1e-6 and abs(got[3].item()) < 1e-6 and got[5].item() == 0.0
    assert abs(got[4].item()) == pytest.approx(pi, abs=2e-6)
    # random angles: same point on the circle, inside (-pi, pi]
    r = (torch.rand(N, generator=torch.Generator().manual_seed(9)) * 2 - 1) * 50.0
    w = port.wrap_to_pi(r)
    assert torch.all(w <= pi + 1e-6) and torch.all(w >= -pi - 1e-6)
    d = np.angle(np.exp(1j * (w.double().numpy() - r.double().numpy())))
    assert np.abs(d).max() < 2e-5


def test_soft_joint_pos_limits_hand_worked():
    """ArticulationData.soft_joint_pos_limits [IL]: mid -+ 0.5 * range * factor, factor 0.9 (assets/unitree.py:53,105,502).
    Go2 URDF limits (go2_description.urdf): hip +-1.0472, front thigh -1.5708 .. 3.4907, rear thigh -0.5236 .. 4.5379,
    calf -2.7227 .. -0.83776; the expected values are worked out by hand."""
    from robot_lab_b200.assets import UNITREE_GO2

    soft = dict(zip(UNITREE_GO2.joint_names, UNITREE_GO2.soft_joint_pos_limits()))
    want = {"FL_hip_joint": (-0.94248, 0.94248), "FR_thigh_joint": (-1.317725, 3.237625),
            "RL_thigh_joint": (-0.270525, 4.284825), "RR_calf_joint": (-2.628453, -0.932007)}
    for name, (lo, hi) in want.items():
        assert soft[name][0] == pytest.approx(lo, abs=2e-6) and soft[name][1] == pytest.approx(hi, abs=2e-6), name


def _dc_table(sat, lim, vlim, kp=20.0, kd=0.5):
    return {"kind": ["dc_motor"], "stiffness": [kp], "damping": [kd], "effort_limit": [lim], "saturation_effort": [sat],
            "velocity_limit": [vlim]}


def test_dcmotor_four_quadrant_hand_worked_cases():
    """DCMotor._clip_effort, the four-quadrant form of IsaacLab >= 2.2 (isaaclab/actuators/actuator_pd.py): the torque
    a DC motor can deliver falls linearly from saturation_effort at zero speed to 0 at velocity_limit (and the braking
    torque rises symmetrically), clipped to +-effort_limit. Go2 leg motor: saturation 23.5, effort limit 23.5,
    velocity limit 30 (assets/unitree.py:107-115). Every expected value is worked out by hand from that line."""
    sat, lim, vlim = 23.5, 23.5, 30.0
    tab = _dc_table(sat, lim, vlim, kp=1.0, kd=0.0)

    def applied(tau_des, vel):
        # kp = 1, kd = 0: desired torque = target - pos
        _, out = port.actuator_step(tab, torch.tensor([[tau_des]]), torch.tensor([[0.0]]), torch.tensor([[vel]]))
        return out.item()

    assert applied(10.0, 0.0) == pytest.approx(10.0)                     # inside the envelope: untouched
    assert applied(40.0, 0.0) == pytest.approx(23.5)                     # zero speed: saturation (= effort limit)
    assert applied(-40.0, 0.0) == pytest.approx(-23.5)
    # driving at half the no-load speed: top line = sat * (1 - 15/30) = 11.75; braking side = sat * (-1 - 0.5) -> -lim
    assert applied(40.0, 15.0) == pytest.approx(11.75)
    assert applied(-40.0, 15.0) == pytest.approx(-23.5)
    # at the no-load speed no driving torque is left; braking is still the full limit
    assert applied(40.0, 30.0) == pytest.approx(0.0, abs=1e-6)
    assert applied(-40.0, 30.0) == pytest.approx(-23.5)
    # beyond it the "driving" bound goes negative: the motor can only brake (top = sat * (1 - 45/30) = -11.75)
    assert applied(40.0, 45.0) == pytest.approx(-11.75)
    # mirrored for negative speeds
    assert applied(-40.0, -15.0) == pytest.approx(-11.75)
    assert applied(40.0, -15.0) == pytest.approx(23.5)
    assert applied(-40.0, -45.0) == pytest.approx(11.75)
    # speeds are clipped at vel_at_effort_lim = vlim * (1 + lim / sat) = 60: the bound stops moving there
    assert applied(40.0, 60.0) == pytest.approx(-23.5)
    assert applied(40.0, 600.0) == pytest.approx(-23.5)


def test_dcmotor_with_effort_limit_below_saturation():
    """A1-style numbers (assets/unitree.py:55-63: saturation 33.5, effort limit 33.5, velocity limit 21) and a
    hypothetical motor whose effort limit sits below saturation: the flat top of the envelope."""
    tab = _dc_table(33.5, 33.5, 21.0, kp=1.0, kd=0.0)
    _, out = port.actuator_step(tab, torch.tensor([[100.0]]), torch.tensor([[0.0]]), torch.tensor([[10.5]]))
    assert out.item() == pytest.approx(33.5 * 0.5)
    tab = _dc_table(40.0, 20.0, 10.0, kp=1.0, kd=0.0)
    for vel, want in ((0.0, 20.0), (2.5, 20.0), (5.0, 20.0), (7.5, 10.0), (10.0, 0.0)):   # 40 * (1 - v/10) capped at 20
        _, out = port.actuator_step(tab, torch.tensor([[100.0]]), torch.tensor([[0.0]]), torch.tensor([[vel]]))
        assert out.item() == pytest.approx(want, abs=1e-5), vel


def test_ideal_pd_torque_is_the_textbook_law():
    tab = {"kind": ["ideal_pd"], "stiffness": [25.0], "damping": [0.5], "effort_limit": [23.5], "saturation_effort": [23.5],
           "velocity_limit": [30.0]}
    comp, app = port.actuator_step(tab, torch.tensor([[0.4]]), torch.tensor([[0.1]]), torch.tensor([[2.0]]))
    assert comp.item() == pytest.approx(25.0 * 0.3 - 0.5 * 2.0)        # kp (q* - q) + kd (0 - qd)
    assert app.item() == pytest.approx(6.5)
    comp, app = port.actuator_step(tab, torch.tensor([[2.0]]), torch.tensor([[0.0]]), torch.tensor([[0.0]]))
    assert comp.item() == pytest.approx(50.0) and app.item() == pytest.approx(23.5)   # clipped to the effort limit


def test_contact_timer_fsm_against_an_event_log():
    """ContactSensor air / contact timers [IL], a second formulation that shares no code with the vectorised restatement:
    per foot, keep the LOG of contact flags per sub-step and derive the four timers from it by definition - current_*_time
    is dt x the length of the current run of (no) contact, last_air_time is dt x (length + 1) of the most recent COMPLETED
    air run (it ended with a touch-down, whose sub-step IsaacLab books to the air time), last_contact_time the same for
    contact runs. Random force sequences around the
    1 N threshold, including exact zeros and exact threshold values."""
    import helpers as H
    cfg, spec = H.make_spec("go2_rough")
    from robot_lab_b200.synthetic import make_state

    n, steps, dt, thr = 64, 40, 0.005, 1.0
    st = make_state(spec, n, seed=11)
    for k in ("current_air_time", "last_air_time", "current_contact_time", "last_contact_time"):
        st[k] = torch.zeros_like(st[k])
    names_h = list(spec.layout.hist_body_names)
    t2h = [names_h.index(nm) for nm in spec.layout.time_body_names]
    g = torch.Generator().manual_seed(5)
    log = []   # per sub-step: [n, feet] bool
    for s_ in range(steps):
        f = torch.randn(n, spec.B, 3, generator=g) * 1.2
        f[torch.rand(n, spec.B, generator=g) < 0.35] = 0.0                  # swing phases: exactly zero force
        f[0, t2h[0]] = torch.tensor([0.6, 0.8, 0.0])                         # |F| == threshold exactly: not a contact (strict >)
        out = port.contact_sensor_update(spec, st, f, dt, thr)
        st = {**st, **out}
        log.append(torch.linalg.vector_norm(f[:, t2h, :], dim=-1) > thr)   # fp32 like the sensor (in double |(0.6f, 0.8f, 0)| > 1)
        flags = torch.stack(log)                                             # [s, n, feet]
        for e in (0, 1, 7, 33, 63):
            for j in range(len(t2h)):
                seq = flags[:, e, j].tolist()
                run, last_air, last_con = 1, 0, 0
                # walk the log: runs of equal flags
                runs = []
                for i in range(1, len(seq)):
                    if seq[i] == seq[i - 1]:
                        run += 1
                    else:
                        runs.append((seq[i - 1], run))
                        run = 1
                cur_flag, cur_run = seq[-1], run
                for fl, ln in runs:                                           # completed runs, oldest first
                    # IsaacLab books the sub-step of the transition to the run that ENDS there
                    # (last_air_time = current_air_time + dt at first contact): a completed run of ln flags lasted ln + 1 steps
                    if fl:
                        last_con = ln + 1
                    else:
                        last_air = ln + 1
                # a run that started at step 0 with the timers at zero: IsaacLab only counts "first contact / first
                # detached" when the previous timer was > 0, which holds for every completed run of length >= 1
                want = {
                    "current_air_time": 0.0 if cur_flag else cur_run * dt,
                    "current_contact_time": cur_run * dt if cur_flag else 0.0,
                    "last_air_time": last_air * dt,
                    "last_contact_time": last_con * dt,
                }
                for k, v in want.items():
                    got = float(st[k][e, j])
                    assert abs(got - v) < 1e-6, f"sub-step {s_} env {e} foot {j} {k}: {got} != {v}"
    # the exact-threshold foot never touched down
    assert float(st["current_contact_time"][0, 0]) == 0.0

"""The IsaacLab-owned term formulas the oracle restates (isaaclab.envs.mdp.rewards / terminations, RewardManager.compute
- not vendored by the reference), re-derived the slow way: plain Python floats, one env and one joint / body at a time,
written from the upstream docstrings ("penalize joint positions if they cross the soft limits", "the max over the
contact-force history", "value = term * weight * dt") rather than from oracle/mdp_port.py's tensor expressions. Catches
slips of the vectorised restatement (axis, clip side, missing abs, wrong accumulation order); it cannot catch a misreading
of upstream that both formulations share - COVERAGE.md keeps calling those terms unpinned."""

import math

import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200.synthetic import make_state

N = 12


def _term(spec, type_name, nth=0):
    return [t for t in spec.rewards if t.type_name == type_name][nth]


def _state(key, seed):
    cfg, spec = H.make_spec(key)
    st = make_state(spec, N, seed=seed)
    return spec, st


def _check(got: torch.Tensor, want: list[float], what: str):
    torch.testing.assert_close(got.double(), torch.tensor(want, dtype=torch.float64), rtol=2e-5, atol=1e-6, msg=lambda m: f"{what}: {m}")


def test_joint_sum_terms_by_hand():
    spec, st = _state("go2_rough", 3)
    for name, field, fn in (("joint_torques_l2", "applied_torque", lambda x: x * x), ("joint_acc_l2", "joint_acc", lambda x: x * x)):
        t = _term(spec, name)
        want = [sum(fn(float(st[field][e, j])) for j in t.joint_ids) for e in range(N)]
        _check(port.reward_term(t, st, spec), want, name)
    t = _term(spec, "action_rate_l2")
    want = [sum((float(st["action"][e, a]) - float(st["prev_action"][e, a])) ** 2 for a in range(spec.A)) for e in range(N)]
    _check(port.reward_term(t, st, spec), want, "action_rate_l2")


def test_joint_deviation_l1_by_hand():
    spec, st = _state("g1_rough", 4)
    for nth in range(3):   # hip / arms / torso groups
        t = _term(spec, "joint_deviation_l1", nth)
        want = [sum(abs(float(st["joint_pos"][e, j]) - spec.default_joint_pos[j]) for j in t.joint_ids) for e in range(N)]
        _check(port.reward_term(t, st, spec), want, f"joint_deviation_l1[{nth}]")


def test_joint_pos_limits_by_hand():
    """'Penalize joint positions if they cross the soft limits': the amount by which q is below the lower or above the
    upper soft limit, summed over the joints."""
    spec, st = _state("go2_rough", 5)
    t = _term(spec, "joint_pos_limits")
    st["joint_pos"][0, t.joint_ids[0]] = spec.soft_pos_limits[t.joint_ids[0]][0] - 0.25      # 0.25 rad below the lower limit
    st["joint_pos"][1, t.joint_ids[1]] = spec.soft_pos_limits[t.joint_ids[1]][1] + 0.125     # 0.125 rad above the upper limit
    st["joint_pos"][2, t.joint_ids[2]] = spec.soft_pos_limits[t.joint_ids[2]][1]              # exactly on the limit: nothing
    want = []
    for e in range(N):
        s = 0.0
        for j in t.joint_ids:
            q, (lo, hi) = float(st["joint_pos"][e, j]), spec.soft_pos_limits[j]
            if q < lo:
                s += lo - q
            if q > hi:
                s += q - hi
        want.append(s)
    _check(port.reward_term(t, st, spec), want, "joint_pos_limits")


def test_contact_force_terms_by_hand():
    """contact_forces [IL]: per body the LARGEST force norm over the history, the excess over the threshold, summed.
    illegal_contact [IL]: any body whose largest force norm over the history exceeds the threshold."""
    spec, st = _state("go2_rough", 6)
    t = _term(spec, "contact_forces")
    h = st["net_forces_w_history"]                      # [N, T, B, 3]
    want = []
    for e in range(N):
        s = 0.0
        for b in t.body_ids:
            worst = max(math.sqrt(sum(float(h[e, k, b, c]) ** 2 for c in range(3))) for k in range(h.shape[1]))
            s += max(worst - t.p[0], 0.0)
        want.append(s)
    _check(port.reward_term(t, st, spec), want, "contact_forces")

    spec, st = _state("g1_rough", 7)
    d = [x for x in spec.dones if x.type_name == "illegal_contact"][0]
    h = st["net_forces_w_history"]
    ep = st["episode_length"] + 1
    got = port.done_term(d, st, spec, ep)
    for e in range(N):
        fired = any(max(math.sqrt(sum(float(h[e, k, b, c]) ** 2 for c in range(3))) for k in range(h.shape[1])) > d.p[0]
                    for b in d.body_ids)
        assert bool(got[e]) == fired, f"illegal_contact env {e}"


def test_time_out_and_bounds_by_hand():
    spec, st = _state("go2_rough", 8)
    st["episode_length"][0] = spec.max_episode_length - 1     # the increment of this step reaches the limit: time out
    st["episode_length"][1] = spec.max_episode_length - 2     # one short
    st["root_pos_w"][2, 0] = 1e3                               # far outside the terrain
    ep, terminated, truncated, bits = port.compute_dones(spec, st)
    names = [d.type_name for d in spec.dones]
    i_to, i_ob = names.index("time_out"), names.index("terrain_out_of_bounds")
    assert bool((bits[0] >> i_to) & 1) and not bool((bits[1] >> i_to) & 1)
    assert bool(truncated[0]) and not bool(terminated[0]), "time_out is a truncation, not a termination"
    assert bool((bits[2] >> i_ob) & 1)
    d = spec.dones[i_ob]
    for e in range(N):
        x, y = float(st["root_pos_w"][e, 0]), float(st["root_pos_w"][e, 1])
        outside = d.p[2] != 0.0 and (abs(x) > d.p[0] or abs(y) > d.p[1])
        assert bool((bits[e] >> i_ob) & 1) == outside, f"terrain_out_of_bounds env {e}"
    assert torch.equal(ep, st["episode_length"] + 1)


def test_reward_manager_arithmetic_by_hand():
    """RewardManager.compute [IL]: for every term with a non-zero weight, value = term * weight * dt; the reward is their
    sum in declared order, the episode sums advance by value, the per-term step reward is value / dt; is_terminated is the
    terminated flag of this step."""
    spec, st = _state("g1_rough", 9)
    ep, terminated, truncated, bits = port.compute_dones(spec, st)
    total, sums, step_reward = port.compute_rewards(spec, st, terminated)
    st2 = dict(st)
    st2["terminated"] = terminated
    d = port.Derived(st2, spec)
    for e in range(N):
        acc = torch.zeros((), dtype=torch.float32)
        for k, t in enumerate(spec.rewards):
            raw = float(terminated[e]) if t.type_name == "is_terminated" else float(port.reward_term(t, st2, spec, d)[e])
            if t.weight == 0.0:
                assert float(step_reward[e, k]) == 0.0
                continue
            value = torch.tensor(raw, dtype=torch.float32) * t.weight * spec.step_dt    # fp32, like the manager
            acc = acc + value
            assert abs(float(sums[e, k]) - (float(st["episode_sums"][e, k]) + float(value))) <= 1e-6 + 1e-5 * abs(float(value))
            assert abs(float(step_reward[e, k]) - float(value) / spec.step_dt) <= 1e-6 + 1e-5 * abs(float(value) / spec.step_dt)
        assert abs(float(total[e]) - float(acc)) <= 1e-6 + 1e-5 * abs(float(acc)), f"reward env {e}"

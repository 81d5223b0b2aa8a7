"""Per-term GPU parity: every reward term evaluated alone through ``rl_term_eval`` (the term-function protocol,
V/mdp/rewards.py:22 ff.) against the oracle's restatement of that term."""

import copy
import math

import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200 import mdp
from robot_lab_b200.cfg import RewardTermCfg, SceneEntityCfg
from robot_lab_b200.spec import compile_reward_term, compile_step_spec
from robot_lab_b200.synthetic import make_state
from robot_lab_b200.tasks import make_env_cfg

pytestmark = pytest.mark.gpu


def _check_terms(spec, terms, n=2048):
    from robot_lab_b200.engine import MdpStepEngine

    st = make_state(spec, n)
    st["terminated"] = torch.rand(n, generator=torch.Generator().manual_seed(3)) < 0.1
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n)
    b.load_logical(st)
    term_dev = st["terminated"].to(torch.uint8).cuda()
    d = port.Derived(st, spec)
    failures = []
    for t in terms:
        got = eng.term_eval(t, b, terminated=term_dev).cpu()
        ref = port.reward_term(t, st, spec, d)
        try:
            torch.testing.assert_close(got, ref, rtol=H.RTOL, atol=H.ATOL)
        except AssertionError as e:
            failures.append(f"{t.name}: {str(e).splitlines()[-3:]}")
    eng.close()
    assert not failures, "\n".join(failures)


@pytest.mark.parametrize("key", ["a1_flat", "go2_rough", "g1_rough", "g1_rough_37"])
def test_active_terms(native_lib, key):
    cfg, spec = H.make_spec(key)
    _check_terms(spec, spec.rewards)


def test_whole_catalogue_including_inactive_terms(native_lib):
    """The terms the in-scope tasks leave at weight 0 (V/velocity_env_cfg.py:379-644) are implemented too."""
    cfg = make_env_cfg(H.TASKS["go2_rough"])
    base = make_env_cfg(H.TASKS["go2_rough"])
    # rebuild the full catalogue with Go2 names filled in, nothing disabled
    from robot_lab_b200.tasks.locomotion_velocity import _reward_catalogue

    cat = _reward_catalogue()
    foot = [".*_foot"]
    for name in ("feet_air_time", "feet_air_time_variance", "feet_contact", "feet_contact_without_cmd", "feet_stumble",
                 "contact_forces"):
        getattr(cat, name).params["sensor_cfg"].body_names = foot
    cat.undesired_contacts.params["sensor_cfg"].body_names = ["^(?!.*_foot).*"]
    cat.feet_slide.params["sensor_cfg"].body_names = foot
    for name in ("feet_slide", "feet_height", "feet_height_body", "feet_distance_y_exp"):
        getattr(cat, name).params["asset_cfg"].body_names = foot
    cat.feet_distance_y_exp.params["stance_width"] = 0.3
    cat.feet_gait.params["synced_feet_pair_names"] = (("FL_foot", "RR_foot"), ("FR_foot", "RL_foot"))
    cat.base_height_l2.params["sensor_cfg"] = None
    cat.base_height_l2.params["target_height"] = 0.33
    cat.joint_mirror.params["mirror_joints"] = [["FR_(hip|thigh|calf).*", "RL_(hip|thigh|calf).*"],
                                                ["FL_(hip|thigh|calf).*", "RR_(hip|thigh|calf).*"]]
    cat.action_mirror.params["mirror_joints"] = cat.joint_mirror.params["mirror_joints"]
    cat.wheel_vel_penalty.params["asset_cfg"].joint_names = [".*_calf_joint"]
    cat.wheel_vel_penalty.params["sensor_cfg"].body_names = foot
    cat.feet_distance_xy_exp = RewardTermCfg(func=mdp.feet_distance_xy_exp, weight=0.0, params={
        "std": math.sqrt(0.25), "asset_cfg": SceneEntityCfg("robot", body_names=foot), "stance_length": 0.4,
        "stance_width": 0.3})
    cat.joint_deviation_hip = RewardTermCfg(func=mdp.joint_deviation_l1, weight=0.0, params={
        "asset_cfg": SceneEntityCfg("robot", joint_names=[".*_hip_joint"])})
    for _, term in cat.items():
        term.weight = 1.0
    cfg.rewards = cat
    layout = cfg.scene.make_layout()  # full layout: every body everywhere
    # time / asset body tensors are limited to 16 bodies in the ABI -> use a feet-only space for them
    from robot_lab_b200.spec import SceneLayout

    feet = tuple(n for n in layout.asset.body_names if n.endswith("_foot"))
    layout = SceneLayout(layout.asset, layout.hist_body_names, feet, feet, layout.terrain, layout.num_rays, layout.hist_len)
    spec = compile_step_spec(cfg, layout)
    assert spec.K >= 34
    _check_terms(spec, spec.rewards)
    del base

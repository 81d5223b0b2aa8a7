"""Spec compiler: the cfg tree resolves to the dimensions, term order, ids and byte counts SURVEY.md section 8 states."""

import math

import pytest

import helpers as H
from robot_lab_b200.cfg import resolve_matching_names, resolve_matching_names_values
from robot_lab_b200.spec import compact_layout, compile_step_spec
from robot_lab_b200.tasks import list_tasks, make_env_cfg

GO2_ORDER = ["lin_vel_z_l2", "ang_vel_xy_l2", "joint_torques_l2", "joint_acc_l2", "joint_pos_limits", "joint_power",
             "stand_still", "joint_pos_penalty", "joint_mirror", "action_rate_l2", "undesired_contacts", "contact_forces",
             "track_lin_vel_xy_exp", "track_ang_vel_z_exp", "feet_air_time", "feet_air_time_variance", "feet_gait",
             "feet_contact_without_cmd", "feet_slide", "feet_height_body", "upward"]


@pytest.mark.parametrize("key,J,B,F,R,K,pol,crit,nbytes", [
    ("a1_flat", 12, 17, 4, 0, 17, 45, 48, 1807),
    ("go2_flat", 12, 19, 4, 0, 21, 45, 48, 1927),
    ("go2_rough", 12, 19, 4, 187, 21, 45, 235, 3423),
    ("g1_rough", 29, 3, 2, 187, 16, 96, 286, 3659),
])
def test_dimensions_and_algorithmic_bytes(key, J, B, F, R, K, pol, crit, nbytes):
    cfg, spec = H.make_spec(key)
    assert (spec.J, spec.B, spec.Bt, spec.R, spec.K) == (J, B, F, R, K)
    assert (spec.obs[0].dim, spec.obs[1].dim) == (pol, crit)
    assert spec.algorithmic_bytes_per_env_step() == nbytes
    assert spec.max_episode_length == 1000 and math.isclose(spec.step_dt, 0.02)


def test_go2_reward_order_weights_and_ids():
    cfg, spec = H.make_spec("go2_rough")
    assert [t.name for t in spec.rewards] == GO2_ORDER
    w = {t.name: t.weight for t in spec.rewards}
    assert (w["lin_vel_z_l2"], w["joint_torques_l2"], w["track_lin_vel_xy_exp"], w["feet_gait"]) == (-2.0, -2.5e-5, 3.0, 0.5)
    t = {t.name: t for t in spec.rewards}
    assert t["joint_mirror"].idx_a == [1, 5, 9, 0, 4, 8] and t["joint_mirror"].idx_b == [2, 6, 10, 3, 7, 11]
    assert t["joint_mirror"].p[0] == 0.5
    assert t["undesired_contacts"].body_ids == list(range(15)) and t["contact_forces"].body_ids == [15, 16, 17, 18]
    assert t["feet_gait"].idx_a == [0, 3, 1, 2]  # (FL, RR), (FR, RL) in the feet-only time space
    assert t["feet_gait"].p[1] == 0.2 ** 2 and t["track_lin_vel_xy_exp"].p[0] == math.sqrt(0.25) ** 2
    assert [d.name for d in spec.dones] == ["time_out", "terrain_out_of_bounds"]
    assert spec.dones[1].p[:3] == [0.5 * (10 * 8 + 40) - 3.0, 0.5 * (20 * 8 + 40) - 3.0, 1.0]  # |x|>57, |y|>97


def test_go2_observation_layout_and_joint_permutation():
    cfg, spec = H.make_spec("go2_rough")
    pol, crit = spec.obs
    assert [t.name for t in pol.terms] == ["base_ang_vel", "projected_gravity", "velocity_commands", "joint_pos", "joint_vel", "actions"]
    assert [t.name for t in crit.terms][-1] == "height_scan" and crit.terms[-1].dim == 187 and crit.terms[-1].clip == (-1.0, 1.0)
    perm = [1, 5, 9, 0, 4, 8, 3, 7, 11, 2, 6, 10]  # cfg order FR,FL,RR,RL x hip,thigh,calf in native BFS ids
    assert pol.terms[3].ids == perm and spec.action.joint_ids == perm
    assert crit.terms[4].ids == list(range(12))  # critic joint terms use the native order
    assert pol.terms[0].scale == 0.25 and pol.terms[4].scale == 0.05 and pol.terms[0].noise == (-0.2, 0.2)
    assert all(t.noise is None for t in crit.terms)  # corruption off for the critic group
    assert spec.action.scale == [0.125, 0.25, 0.25] * 4 and spec.action.clip[0] == (-100.0, 100.0)


def test_g1_specifics():
    cfg, spec = H.make_spec("g1_rough")
    names = [t.name for t in spec.rewards]
    assert names[0] == "is_terminated" and names[-3:] == ["joint_deviation_hip_l1", "joint_deviation_arms_l1", "joint_deviation_torso_l1"]
    t = {t.name: t for t in spec.rewards}
    assert t["track_lin_vel_xy_exp"].type_name == "track_lin_vel_xy_yaw_frame_exp"
    assert t["feet_air_time"].type_name == "feet_air_time_positive_biped" and t["feet_air_time"].p[0] == 0.4
    assert len(t["joint_torques_l2"].joint_ids) == 12 and len(t["joint_acc_l2"].joint_ids) == 8
    assert spec.layout.hist_body_names == ("torso_link", "left_ankle_roll_link", "right_ankle_roll_link")
    assert [d.name for d in spec.dones][-1] == "illegal_contact" and spec.dones[-1].body_ids == [0]
    j = spec.layout.asset.joint_names.index("left_knee_joint")
    assert math.isclose(spec.action.scale[spec.action.joint_ids.index(j)], 0.25 * 139.0 / (0.025101925 * (20 * 3.1415926535) ** 2))


def test_every_registered_task_compiles_and_fits_the_abi():
    for task in list_tasks():
        cfg = make_env_cfg(task)
        for layout in (compact_layout(cfg), cfg.scene.make_layout()):
            spec = compile_step_spec(make_env_cfg(task), layout)
            c = spec.to_ctypes()
            assert c.num_reward_terms == spec.K and c.obs[1].dim == spec.obs[1].dim


def test_name_resolution_semantics():
    names = ["FL_hip", "FR_hip", "FL_thigh", "FR_thigh"]
    assert resolve_matching_names(["FR_.*", "FL_.*"], names)[0] == [0, 1, 2, 3]  # target order
    assert resolve_matching_names(["FR_.*", "FL_.*"], names, preserve_order=True)[0] == [1, 3, 0, 2]  # key order
    with pytest.raises(ValueError):
        resolve_matching_names(["nope"], names)
    with pytest.raises(ValueError):
        resolve_matching_names([".*_hip", "FL_.*"], names)  # FL_hip matched twice
    ids, _, vals = resolve_matching_names_values({".*_hip": 0.125, "^(?!.*_hip).*": 0.25}, names)
    assert ids == [0, 1, 2, 3] and vals == [0.125, 0.125, 0.25, 0.25]


def test_unsupported_configurations_are_rejected():
    from robot_lab_b200.cfg import SceneEntityCfg

    cfg = make_env_cfg(H.TASKS["go2_rough"])
    cfg2 = make_env_cfg(H.TASKS["go2_rough"])
    from robot_lab_b200.tasks.locomotion_velocity import _reward_catalogue

    cat = _reward_catalogue()
    cat.base_height_l2.weight = 1.0  # still carries the ray-sensor cfg -> batch-global branch (rewards.py:634)
    for name, term in cat.items():
        if name != "base_height_l2":
            setattr(cat, name, None)
    cfg2.rewards = cat
    with pytest.raises(NotImplementedError):
        compile_step_spec(cfg2, cfg2.scene.make_layout())
    cfg.rewards.feet_gait.params["synced_feet_pair_names"] = (("FL_foot",), ("FR_foot", "RL_foot"))
    with pytest.raises(ValueError, match="two pairs"):
        compile_step_spec(cfg, cfg.scene.make_layout())
    del SceneEntityCfg


def test_per_launch_byte_accounting():
    """the fused figure of SURVEY.md 8(d) and its split over the two launches of an env step."""
    import helpers as H

    _, spec = H.make_spec("go2_rough")
    assert spec.algorithmic_bytes_per_env_step() == 3423
    pre, post = spec.algorithmic_bytes_per_launch("pre_reset"), spec.algorithmic_bytes_per_launch("post_reset")
    assert (pre, post) == (1463, 2126)
    # both launches read root state (13 words), joint pos/vel (2J) and the command (3): the split exceeds the fused
    # figure by those re-reads plus command state the fused accounting keeps on chip, minus the joint target /
    # stored action that process_action owns
    assert 3423 < pre + post < 3423 + 4 * (13 + 2 * spec.J + 3 + 20)


def _two_term_action_cfg():
    """Go2 with the calf joints driven like the wheels of the wheeled robots: a JointPositionAction term for hips and
    thighs and a JointVelocityAction term for the rest (V/config/wheeled/unitree_go2w/rough_env_cfg.py:22-32,101-106)."""
    from robot_lab_b200.cfg import JointPositionActionCfg, JointVelocityActionCfg
    from robot_lab_b200.tasks import make_env_cfg

    cfg = make_env_cfg(H.TASKS["go2_rough"])
    legs = [f"{leg}_{part}_joint" for leg in ("FR", "FL", "RR", "RL") for part in ("hip", "thigh")]
    wheels = [f"{leg}_calf_joint" for leg in ("FR", "FL", "RR", "RL")]
    cfg.actions.joint_pos = JointPositionActionCfg(joint_names=legs, scale={".*_hip_joint": 0.125, "^(?!.*_hip_joint).*": 0.25},
                                                   use_default_offset=True, clip={".*": (-100.0, 100.0)}, preserve_order=True)
    cfg.actions.joint_vel = JointVelocityActionCfg(joint_names=wheels, scale=5.0, use_default_offset=True,
                                                   clip={".*": (-100.0, 100.0)}, preserve_order=True)
    return cfg, legs, wheels


def test_two_action_terms_concatenate_in_declaration_order():
    from robot_lab_b200.spec import compact_layout, compile_step_spec

    cfg, legs, wheels = _two_term_action_cfg()
    spec = compile_step_spec(cfg, compact_layout(cfg))
    a = spec.action
    assert spec.A == 12 and a.joint_names == legs + wheels
    assert a.kind == [0] * 8 + [1] * 4
    names = list(spec.layout.asset.joint_names)
    assert a.joint_ids == [names.index(n) for n in legs + wheels]
    assert a.scale[:2] == [0.125, 0.25] and a.scale[8:] == [5.0] * 4
    dj = spec.layout.asset.default_joint_pos()
    assert a.offset[:8] == [dj[j] for j in a.joint_ids[:8]] and a.offset[8:] == [0.0] * 4   # default joint velocity
    c = spec.to_ctypes().action
    assert list(c.target_kind[:12]) == a.kind and c.n_actions == 12
    # last_action / action_rate_l2 see the full 12-wide action vector
    assert spec.obs[0].terms[-1].type_name == "last_action" and spec.obs[0].terms[-1].dim == 12

"""Per-robot task cfgs for the in-scope robots: Unitree A1, Go2 (quadrupeds) and G1 (humanoid).

Values follow the reference override chains:
  * Go2 : V/config/quadruped/unitree_go2/rough_env_cfg.py:18-161, flat_env_cfg.py:10-29
  * A1  : V/config/quadruped/unitree_a1/rough_env_cfg.py:18-159,  flat_env_cfg.py:10-29
  * G1  : V/config/humanoid/unitree_g1/rough_env_cfg.py:16-167,   flat_env_cfg.py:10-37
Only the MDP-step part is carried over (observations, actions, rewards, terminations, commands); event and
curriculum overrides belong to the physics / host side.
"""

from __future__ import annotations

from .. import mdp
from ..assets import UNITREE_A1, UNITREE_G1_29DOF, UNITREE_G1_37DOF, UNITREE_GO2, RobotAsset
from .locomotion_velocity import LocomotionVelocityRoughEnvCfg

_QUAD_JOINT_ORDER = [f"{leg}_{part}_joint" for leg in ("FR", "FL", "RR", "RL") for part in ("hip", "thigh", "calf")]


class _QuadrupedRoughEnvCfg(LocomotionVelocityRoughEnvCfg):
    """Shared body of the A1 / Go2 rough cfgs (the two reference files differ in 7 values)."""

    base_link_name = "base"
    foot_link_name = ".*_foot"
    joint_names = _QUAD_JOINT_ORDER
    _asset: RobotAsset = UNITREE_GO2
    # (feet_air_time, feet_air_time_variance, feet_slide, feet_gait, base_height target)
    _weights = {"feet_air_time": 0.1, "feet_air_time_variance": -1.0, "feet_slide": -0.1, "feet_gait": 0.5}
    _base_height_target = 0.33

    def __init__(self) -> None:
        super().__init__()
        self.scene.robot = self._asset
        self.scene.foot_body_regex = self.foot_link_name
        # observations
        pol = self.observations.policy
        pol.base_lin_vel.scale = 2.0
        pol.base_ang_vel.scale = 0.25
        pol.joint_pos.scale = 1.0
        pol.joint_vel.scale = 0.05
        pol.base_lin_vel = None
        pol.height_scan = None
        pol.joint_pos.params["asset_cfg"].joint_names = self.joint_names
        pol.joint_vel.params["asset_cfg"].joint_names = self.joint_names
        # actions
        act = self.actions.joint_pos
        act.scale = {".*_hip_joint": 0.125, "^(?!.*_hip_joint).*": 0.25}
        act.clip = {".*": (-100.0, 100.0)}
        act.joint_names = self.joint_names
        # rewards
        r, w = self.rewards, self._weights
        foot, not_foot = [self.foot_link_name], [f"^(?!.*{self.foot_link_name}).*"]
        r.is_terminated.weight = 0
        r.lin_vel_z_l2.weight = -2.0
        r.ang_vel_xy_l2.weight = -0.05
        r.flat_orientation_l2.weight = 0
        r.base_height_l2.weight = 0
        r.base_height_l2.params["target_height"] = self._base_height_target
        r.base_height_l2.params["asset_cfg"].body_names = [self.base_link_name]
        r.joint_torques_l2.weight = -2.5e-5
        r.joint_vel_l2.weight = 0
        r.joint_acc_l2.weight = -2.5e-7
        r.joint_pos_limits.weight = -5.0
        r.joint_vel_limits.weight = 0
        r.joint_power.weight = -2e-5
        r.stand_still.weight = -2.0
        r.joint_pos_penalty.weight = -1.0
        r.joint_mirror.weight = -0.05
        r.joint_mirror.params["mirror_joints"] = [
            ["FR_(hip|thigh|calf).*", "RL_(hip|thigh|calf).*"],
            ["FL_(hip|thigh|calf).*", "RR_(hip|thigh|calf).*"],
        ]
        r.action_rate_l2.weight = -0.01
        r.undesired_contacts.weight = -1.0
        r.undesired_contacts.params["sensor_cfg"].body_names = not_foot
        r.contact_forces.weight = -1.5e-4
        r.contact_forces.params["sensor_cfg"].body_names = foot
        r.track_lin_vel_xy_exp.weight = 3.0
        r.track_ang_vel_z_exp.weight = 1.5
        r.feet_air_time.weight = w["feet_air_time"]
        r.feet_air_time.params["threshold"] = 0.5
        r.feet_air_time.params["sensor_cfg"].body_names = foot
        r.feet_air_time_variance.weight = w["feet_air_time_variance"]
        r.feet_air_time_variance.params["sensor_cfg"].body_names = foot
        r.feet_contact.weight = 0
        r.feet_contact.params["sensor_cfg"].body_names = foot
        r.feet_contact_without_cmd.weight = 0.1
        r.feet_contact_without_cmd.params["sensor_cfg"].body_names = foot
        r.feet_stumble.weight = 0
        r.feet_stumble.params["sensor_cfg"].body_names = foot
        r.feet_slide.weight = w["feet_slide"]
        r.feet_slide.params["sensor_cfg"].body_names = foot
        r.feet_slide.params["asset_cfg"].body_names = foot
        r.feet_height.weight = 0
        r.feet_height.params["target_height"] = 0.05
        r.feet_height.params["asset_cfg"].body_names = foot
        r.feet_height_body.weight = -5.0
        r.feet_height_body.params["target_height"] = -0.2
        r.feet_height_body.params["asset_cfg"].body_names = foot
        r.feet_gait.weight = w["feet_gait"]
        r.feet_gait.params["synced_feet_pair_names"] = (("FL_foot", "RR_foot"), ("FR_foot", "RL_foot"))
        r.upward.weight = 1.0
        if type(self)._is_leaf_rough:
            self.disable_zero_weight_rewards()
        # terminations
        self.terminations.illegal_contact = None

    _is_leaf_rough = True


class UnitreeGo2RoughEnvCfg(_QuadrupedRoughEnvCfg):
    task_name = "RobotLab-Isaac-Velocity-Rough-Unitree-Go2-v0"


class UnitreeA1RoughEnvCfg(_QuadrupedRoughEnvCfg):
    task_name = "RobotLab-Isaac-Velocity-Rough-Unitree-A1-v0"
    _asset = UNITREE_A1
    _weights = {"feet_air_time": 0, "feet_air_time_variance": 0, "feet_slide": 0, "feet_gait": 0}
    _base_height_target = 0.35


def _flatten(cfg: LocomotionVelocityRoughEnvCfg) -> None:
    """The common body of every flat_env_cfg.py: plane terrain, no height scanner (GO2/flat_env_cfg.py:16-25)."""
    cfg.rewards.base_height_l2 and cfg.rewards.base_height_l2.params.__setitem__("sensor_cfg", None)
    cfg.scene.terrain.terrain_type = "plane"
    cfg.scene.height_scanner = None
    cfg.observations.policy.height_scan = None
    cfg.observations.critic.height_scan = None


class UnitreeGo2FlatEnvCfg(UnitreeGo2RoughEnvCfg):
    task_name = "RobotLab-Isaac-Velocity-Flat-Unitree-Go2-v0"
    _is_leaf_rough = False

    def __init__(self) -> None:
        super().__init__()
        _flatten(self)
        self.disable_zero_weight_rewards()


class UnitreeA1FlatEnvCfg(UnitreeA1RoughEnvCfg):
    task_name = "RobotLab-Isaac-Velocity-Flat-Unitree-A1-v0"
    _is_leaf_rough = False

    def __init__(self) -> None:
        super().__init__()
        _flatten(self)
        self.disable_zero_weight_rewards()


class UnitreeG1RoughEnvCfg(LocomotionVelocityRoughEnvCfg):
    task_name = "RobotLab-Isaac-Velocity-Rough-Unitree-G1-v0"
    base_link_name = "torso_link"
    foot_link_name = ".*_ankle_roll_link"
    _asset: RobotAsset = UNITREE_G1_29DOF
    _is_leaf_rough = True

    def __init__(self) -> None:
        super().__init__()
        self.scene.robot = self._asset
        self.scene.foot_body_regex = self.foot_link_name
        pol = self.observations.policy
        pol.base_lin_vel.scale = 2.0
        pol.base_ang_vel.scale = 0.25
        pol.joint_pos.scale = 1.0
        pol.joint_vel.scale = 0.05
        pol.base_lin_vel = None
        pol.height_scan = None
        act = self.actions.joint_pos
        act.scale = dict(self._asset.action_scale_patterns)
        act.clip = {".*": (-100.0, 100.0)}
        r = self.rewards
        foot, not_foot = [self.foot_link_name], [f"^(?!.*{self.foot_link_name}).*"]
        leg_mirror = [["left_(hip|knee|ankle).*", "right_(hip|knee|ankle).*"]]
        r.is_terminated.weight = -200.0
        r.lin_vel_z_l2.weight = 0
        r.ang_vel_xy_l2.weight = -0.1
        r.flat_orientation_l2.weight = -0.2
        r.base_height_l2.weight = 0
        r.base_height_l2.params["target_height"] = 0
        r.base_height_l2.params["asset_cfg"].body_names = [self.base_link_name]
        r.joint_torques_l2.weight = -1.5e-7
        r.joint_torques_l2.params["asset_cfg"].joint_names = [".*_hip_.*", ".*_knee_joint", ".*_ankle_.*"]
        r.joint_vel_l2.weight = 0
        r.joint_acc_l2.weight = -1.25e-7
        r.joint_acc_l2.params["asset_cfg"].joint_names = [".*_hip_.*", ".*_knee_joint"]
        self.create_joint_deviation_l1_rewterm("joint_deviation_hip_l1", -0.1, [".*hip_yaw.*", ".*hip_roll.*"])
        self.create_joint_deviation_l1_rewterm("joint_deviation_arms_l1", -0.1, [".*shoulder.*", ".*elbow.*"])
        self.create_joint_deviation_l1_rewterm("joint_deviation_torso_l1", -0.1, ["waist_yaw_joint"])
        r.joint_pos_limits.weight = -0.5
        r.joint_vel_limits.weight = 0
        r.joint_power.weight = 0
        r.stand_still.weight = 0
        r.joint_pos_penalty.weight = -1.0
        r.joint_mirror.weight = 0
        r.joint_mirror.params["mirror_joints"] = leg_mirror
        r.action_rate_l2.weight = -0.005
        r.action_mirror.weight = 0
        r.action_mirror.params["mirror_joints"] = leg_mirror
        r.undesired_contacts.weight = 0
        r.undesired_contacts.params["sensor_cfg"].body_names = not_foot
        r.contact_forces.weight = 0
        r.contact_forces.params["sensor_cfg"].body_names = foot
        r.track_lin_vel_xy_exp.weight = 3.0
        r.track_lin_vel_xy_exp.func = mdp.track_lin_vel_xy_yaw_frame_exp
        r.track_ang_vel_z_exp.weight = 3.0
        r.track_ang_vel_z_exp.func = mdp.track_ang_vel_z_world_exp
        r.feet_air_time.weight = 0.25
        r.feet_air_time.func = mdp.feet_air_time_positive_biped
        r.feet_air_time.params["threshold"] = 0.4
        r.feet_air_time.params["sensor_cfg"].body_names = foot
        r.feet_contact.weight = 0
        r.feet_contact.params["sensor_cfg"].body_names = foot
        r.feet_contact_without_cmd.weight = 0
        r.feet_contact_without_cmd.params["sensor_cfg"].body_names = foot
        r.feet_stumble.weight = 0
        r.feet_stumble.params["sensor_cfg"].body_names = foot
        r.feet_slide.weight = -0.2
        r.feet_slide.params["sensor_cfg"].body_names = foot
        r.feet_slide.params["asset_cfg"].body_names = foot
        r.feet_height.weight = 0
        r.feet_height.params["target_height"] = 0.05
        r.feet_height.params["asset_cfg"].body_names = foot
        r.feet_height_body.weight = 0
        r.feet_height_body.params["target_height"] = -0.2
        r.feet_height_body.params["asset_cfg"].body_names = foot
        r.upward.weight = 1.0
        if type(self)._is_leaf_rough:
            self.disable_zero_weight_rewards()
        self.terminations.illegal_contact.params["sensor_cfg"].body_names = [self.base_link_name]
        rng = self.commands.base_velocity.ranges
        rng.lin_vel_x, rng.lin_vel_y, rng.ang_vel_z = (-1.0, 1.0), (-1.0, 1.0), (-1.0, 1.0)


class UnitreeG1FlatEnvCfg(UnitreeG1RoughEnvCfg):
    task_name = "RobotLab-Isaac-Velocity-Flat-Unitree-G1-v0"
    _is_leaf_rough = False

    def __init__(self) -> None:
        super().__init__()
        _flatten(self)
        r = self.rewards
        r.track_ang_vel_z_exp.weight = 1.0
        r.lin_vel_z_l2.weight = -0.2
        r.action_rate_l2.weight = -0.005
        r.joint_acc_l2.weight = -1.0e-7
        r.joint_torques_l2.weight = -2.0e-6
        r.joint_torques_l2.params["asset_cfg"].joint_names = [".*_hip_.*", ".*_knee_joint"]
        self.disable_zero_weight_rewards()


class UnitreeG1Rough37DofEnvCfg(UnitreeG1RoughEnvCfg):
    """Synthetic J = 37 variant (BASELINE.json labels config 4 "37 DoF"; the reference's G1 has 29)."""

    task_name = "RobotLab-Isaac-Velocity-Rough-Unitree-G1-37dof-v0"
    _asset = UNITREE_G1_37DOF

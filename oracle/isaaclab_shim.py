"""Minimal ``isaaclab`` stand-in so that the reference's term files load UNMODIFIED, by path, on CPU.

TEST INFRASTRUCTURE, build-container only: ``/root/reference`` does not exist on the GPU box, so nothing that runs
there imports this module. It is used (a) by ``tests/test_oracle_vs_reference.py`` to pin ``oracle/mdp_port.py``
against the reference's own functions and (b) by ``tests/golden/make_golden.py`` to generate the committed
fixtures. No reference source is copied: the files are executed where they lie.

The reference imports at module top (V/mdp/rewards.py:10-16): ``isaaclab.utils.math`` (quat_apply,
quat_apply_inverse, quat_conjugate, yaw_quat), ``isaaclab.assets``, ``isaaclab.envs.mdp`` (joint_deviation_l1,
used at :101), ``isaaclab.managers`` (ManagerTermBase, SceneEntityCfg, RewardTermCfg), ``isaaclab.sensors``.
The math helpers are restated from IsaacLab v2.3.2 - they are [IL] code, unpinned.
"""

from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import torch

REFERENCE_ROOT = Path("/root/reference")
MDP_DIR = REFERENCE_ROOT / "source/robot_lab/robot_lab/tasks/manager_based/locomotion/velocity/mdp"


def reference_available() -> bool:
    return (MDP_DIR / "rewards.py").exists()


def _math_module() -> types.ModuleType:
    from . import mdp_port as port

    m = types.ModuleType("isaaclab.utils.math")

    def quat_conjugate(q: torch.Tensor) -> torch.Tensor:
        shape = q.shape
        q = q.reshape(-1, 4)
        return torch.cat((q[..., 0:1], -q[..., 1:]), dim=-1).view(shape)

    def _flat(fn):
        def wrapped(quat, vec):
            shape = vec.shape
            return fn(quat.reshape(-1, 4), vec.reshape(-1, 3)).view(shape)
        return wrapped

    m.quat_apply = _flat(port.quat_apply)
    m.quat_apply_inverse = _flat(port.quat_apply_inverse)
    m.quat_conjugate = quat_conjugate
    m.yaw_quat = lambda q: port.yaw_quat(q.reshape(-1, 4)).view(q.shape)
    m.wrap_to_pi = port.wrap_to_pi
    m.quat_from_euler_xyz = port.quat_from_euler_xyz   # [IL], restated
    m.quat_mul = port.quat_mul                         # [IL], restated

    # sample_uniform [IL] = torch.rand(size) * (upper - lower) + lower; the harness feeds the uniforms so that the
    # reference and the port see the same draws (ref_harness.reference_reset_root_state)
    m._uniform_queue = []

    def sample_uniform(lower, upper, size, device=None):
        u = m._uniform_queue.pop(0) if m._uniform_queue else torch.rand(size)
        assert tuple(u.shape) == tuple(size), (u.shape, size)
        return u * (upper - lower) + lower

    m.sample_uniform = sample_uniform
    return m


class SceneEntityCfg:
    def __init__(self, name, joint_names=None, body_names=None, joint_ids=slice(None), body_ids=slice(None),
                 preserve_order=False):
        self.name, self.joint_names, self.body_names = name, joint_names, body_names
        self.joint_ids, self.body_ids, self.preserve_order = joint_ids, body_ids, preserve_order


class ManagerTermBase:
    def __init__(self, cfg, env):
        self.cfg, self._env = cfg, env

    @property
    def num_envs(self):
        return self._env.num_envs

    @property
    def device(self):
        return self._env.device


class RewardTermCfg:
    def __init__(self, func=None, weight=0.0, params=None):
        self.func, self.weight, self.params = func, weight, params or {}


def install() -> None:
    """Register the fake ``isaaclab`` package tree in ``sys.modules`` (idempotent)."""
    if "isaaclab" in sys.modules and getattr(sys.modules["isaaclab"], "__rl_shim__", False):
        return

    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    root = mod("isaaclab")
    root.__rl_shim__ = True
    root.__path__ = []
    utils = mod("isaaclab.utils")
    utils.__path__ = []
    math_mod = _math_module()
    sys.modules["isaaclab.utils.math"] = math_mod
    utils.math = math_mod
    assets = mod("isaaclab.assets")
    assets.Articulation = object
    assets.RigidObject = object
    sensors = mod("isaaclab.sensors")
    sensors.ContactSensor = object
    sensors.RayCaster = object
    managers = mod("isaaclab.managers")
    managers.ManagerTermBase = ManagerTermBase
    managers.SceneEntityCfg = SceneEntityCfg
    managers.RewardTermCfg = RewardTermCfg
    envs = mod("isaaclab.envs")
    envs.__path__ = []
    envs.ManagerBasedRLEnv = object
    envs_mdp = mod("isaaclab.envs.mdp")
    envs.mdp = envs_mdp

    def joint_deviation_l1(env, asset_cfg=SceneEntityCfg("robot")):  # [IL] rewards.joint_deviation_l1
        asset = env.scene[asset_cfg.name]
        angle = asset.data.joint_pos[:, asset_cfg.joint_ids] - asset.data.default_joint_pos[:, asset_cfg.joint_ids]
        return torch.sum(torch.abs(angle), dim=1)

    envs_mdp.joint_deviation_l1 = joint_deviation_l1
    root.utils, root.assets, root.sensors, root.managers, root.envs = utils, assets, sensors, managers, envs


def load_reference_module(filename: str):
    """Execute ``V/mdp/<filename>`` from /root/reference as a stand-alone module."""
    if not reference_available():
        raise FileNotFoundError(f"{MDP_DIR} is not present (the reference only exists in the build container)")
    install()
    path = MDP_DIR / filename
    name = f"_reference_mdp_{path.stem}"
    spec = importlib.util.spec_from_file_location(name, path)
    module = importlib.util.module_from_spec(spec)
    sys.modules[name] = module
    spec.loader.exec_module(module)
    return module

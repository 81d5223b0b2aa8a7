"""Task registry: gym-style ids -> env cfg classes (mirrors the ``gym.register`` calls at
V/config/quadruped/unitree_go2/__init__.py:12-32 and siblings; ``gymnasium`` itself is optional)."""

from __future__ import annotations

from .locomotion_velocity import LocomotionVelocityRoughEnvCfg, SceneCfg, SimCfg
from .unitree import (
    UnitreeA1FlatEnvCfg,
    UnitreeA1RoughEnvCfg,
    UnitreeG1FlatEnvCfg,
    UnitreeG1Rough37DofEnvCfg,
    UnitreeG1RoughEnvCfg,
    UnitreeGo2FlatEnvCfg,
    UnitreeGo2RoughEnvCfg,
)

TASKS = {
    c.task_name: c
    for c in (
        UnitreeA1FlatEnvCfg, UnitreeA1RoughEnvCfg, UnitreeGo2FlatEnvCfg, UnitreeGo2RoughEnvCfg,
        UnitreeG1FlatEnvCfg, UnitreeG1RoughEnvCfg, UnitreeG1Rough37DofEnvCfg,
    )
}


def list_tasks() -> list[str]:
    return sorted(TASKS)


def make_env_cfg(task: str, num_envs: int | None = None) -> LocomotionVelocityRoughEnvCfg:
    if task not in TASKS:
        raise KeyError(f"unknown task '{task}'; known: {list_tasks()}")
    cfg = TASKS[task]()
    if num_envs is not None:
        cfg.scene.num_envs = int(num_envs)
    return cfg

"""Rollout storage and the multi-GPU hand-off to PPO.

Envs shard by contiguous env-id ranges, one process per GPU; the MDP step itself needs no collective
(SURVEY.md 8(e)). The one exchange on the path is here: at the end of a rollout every rank all-gathers its
``[steps, N_local, width]`` buffer so that the learner sees the global batch (NCCL over NVLink 5 / NVSwitch through
``torch.distributed``; ``gloo`` in the CPU tests). rsl_rl's own alternative - keep rollouts local and all-reduce
gradients - needs nothing from this module.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from .spec import StepSpec


def rollout_row_width(spec: StepSpec) -> int:
    """fp32 words per (step, env): policy obs, critic obs, action, reward, done, value, log-prob, action mean, sigma."""
    return spec.obs[0].dim + spec.obs[1].dim + spec.A + 4 + 2 * spec.A


def shard_range(num_envs_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous env-id block of a rank; the first ``num_envs_total % world`` ranks take one extra env."""
    base, extra = divmod(num_envs_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class RolloutBuffer:
    """``[steps, N_local, width]`` fp32, filled step by step straight from the step outputs."""

    def __init__(self, spec: StepSpec, num_envs: int, steps: int, device: torch.device | str):
        self.spec, self.N, self.steps = spec, num_envs, steps
        self.width = rollout_row_width(spec)
        self.data = torch.zeros(steps, num_envs, self.width, device=device)
        dp, dc, a = spec.obs[0].dim, spec.obs[1].dim, spec.A
        o = 0
        self.slices = {}
        for name, w in (("obs_policy", dp), ("obs_critic", dc), ("action", a), ("reward", 1), ("done", 1),
                        ("value", 1), ("log_prob", 1), ("mean", a), ("sigma", a)):
            self.slices[name] = slice(o, o + w)
            o += w
        self.t = 0

    def add(self, obs_policy, obs_critic, action, reward, done, value=None, log_prob=None, mean=None, sigma=None):
        row = self.data[self.t]
        row[:, self.slices["obs_policy"]] = obs_policy
        row[:, self.slices["obs_critic"]] = obs_critic
        row[:, self.slices["action"]] = action
        row[:, self.slices["reward"]] = reward.unsqueeze(-1)
        row[:, self.slices["done"]] = done.to(row.dtype).unsqueeze(-1)
        for name, v in (("value", value), ("log_prob", log_prob)):
            if v is not None:
                row[:, self.slices[name]] = v.reshape(self.N, 1)
        for name, v in (("mean", mean), ("sigma", sigma)):
            if v is not None:
                row[:, self.slices[name]] = v
        self.t = (self.t + 1) % self.steps

    def field(self, name: str) -> torch.Tensor:
        return self.data[:, :, self.slices[name]]

    def all_gather(self, group=None) -> torch.Tensor:
        """Global rollout ``[steps, N_total, width]`` (rank-major env order == global env ids)."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return self.data
        world = dist.get_world_size(group)
        # concatenated-along-dim-0 output: the form both NCCL and gloo accept
        out = torch.empty(world * self.steps, self.N, self.width, device=self.data.device, dtype=self.data.dtype)
        dist.all_gather_into_tensor(out, self.data.contiguous(), group=group)
        out = out.view(world, self.steps, self.N, self.width)
        return out.permute(1, 0, 2, 3).reshape(self.steps, world * self.N, self.width)

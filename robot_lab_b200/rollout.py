"""Rollout storage and the multi-GPU hand-off to PPO.

Envs shard by contiguous env-id ranges, one process per GPU; the MDP step itself needs no collective
(SURVEY.md 8(e)). The one exchange on the path is here (reference recipe: README.md:323-337,
scripts/reinforcement_learning/rsl_rl/train.py:143-150): every rank all-gathers its rollout so that the learner sees
the global batch (NCCL over NVLink 5 / NVSwitch through ``torch.distributed``; ``gloo`` in the CPU tests).

Two things keep that exchange off the critical path of the rollout:

* **no copy into the rollout**: a step's results are *written* there. ``RolloutStorage`` owns one contiguous slab per
  step - policy rows, critic rows, reward, done masks (+ the policy's own outputs) - and ``bind(buffers, t)`` points
  the step kernels' outputs at slab ``t`` (observation rows take any pitch, everything else is a plain ``[N]`` plane),
  so ``RolloutBuffer.add``'s nine slice copies per step do not exist;
* **streamed gather**: ``gather_step(t)`` issues the all-gather of slab ``t`` on a side stream as soon as step ``t`` has
  finished, while step ``t + 1`` runs; ``finish()`` joins. What stays exposed is what the links cannot hide: every rank
  has to *receive* ``(world - 1) x steps x N x width`` bytes - 881 MB for 8 x 4096 envs x 24 steps - through its own
  NVLink ingress.

rsl_rl's own alternative - keep rollouts local and all-reduce gradients - needs nothing from
this module; ``bench.py`` measures it beside the gather.
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


class RolloutStorage:
    """``steps`` slabs of ``N * width`` fp32 words each; inside a slab every quantity is its own contiguous plane:

        obs_policy [N, Dp] | obs_critic [N, Dc] | action [N, A] | mean [N, A] | sigma [N, A] | reward [N] | value [N] |
        log_prob [N] | done [N] (fp32 view of: terminated u8 [N], truncated u8 [N], 2 N spare bytes)

    The step kernels write obs_policy / obs_critic / reward / terminated / truncated of step ``t`` straight into slab
    ``t`` (``bind``); the policy side fills action / mean / sigma / value / log_prob. ``gathered`` (allocated when the
    process group has more than one rank) is ``[steps, world, N * width]``: slab ``t`` of every rank, rank-major, so the
    global env id of row ``e`` of rank ``r`` is ``r * N + e``; it is stored by chunks of ``gather_every`` steps so that every
    all-gather writes its contiguous output in place (no staging copy)."""

    def __init__(self, spec: StepSpec, num_envs: int, steps: int, device: torch.device | str, group=None, gather_every: int = 1):
        self.spec, self.N, self.steps, self.group = spec, int(num_envs), int(steps), group
        self.every = max(1, int(gather_every))      # slabs per all-gather: fewer, larger collectives (launch latency vs overlap)
        if self.steps % self.every:
            raise ValueError("gather_every must divide the number of steps")
        self.width = rollout_row_width(spec)
        dev = torch.device(device)
        self.data = torch.zeros(self.steps, self.N * self.width, device=dev)
        N, dp, dc, a = self.N, spec.obs[0].dim, spec.obs[1].dim, spec.A
        self._planes, o = {}, 0
        for name, w in (("obs_policy", dp), ("obs_critic", dc), ("action", a), ("mean", a), ("sigma", a), ("reward", 1),
                        ("value", 1), ("log_prob", 1), ("done", 1)):
            self._planes[name] = (o, w)
            o += N * w
        assert o == N * self.width
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        # [chunks, world, every * slab]: chunk c of every rank, rank-major - exactly what an all-gather of data[c*every:(c+1)*every] produces
        self._gath = (torch.empty(self.steps // self.every, self.world, self.every * self.N * self.width, device=dev)
                      if self.world > 1 else None)
        self._side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._events = [torch.cuda.Event() for _ in range(self.steps)] if dev.type == "cuda" else None
        self._whole = None   # [world, steps * N * width], only for the un-streamed baseline (gather_all)

    def set_gather_every(self, k: int) -> None:
        """Change the chunking of the streamed gather (re-allocates the gathered buffer; the slabs stay where they are)."""
        k = max(1, int(k))
        if self.steps % k:
            raise ValueError("gather_every must divide the number of steps")
        self.every = k
        if self.world > 1:
            self._gath = torch.empty(self.steps // k, self.world, k * self.N * self.width, device=self.data.device)

    @property
    def gathered(self) -> torch.Tensor | None:
        """``[steps, world, N * width]`` view of the gathered rollout (slab ``t`` of rank ``r`` at ``[t, r]``)."""
        if self._gath is None:
            return None
        c, w = self._gath.shape[0], self.world
        return self._gath.view(c, w, self.every, self.N * self.width).permute(0, 2, 1, 3).reshape(self.steps, w, self.N * self.width) \
            if self.every > 1 else self._gath

    # ---- views -----------------------------------------------------------------------------------------------------
    def plane(self, name: str, t: int, data: torch.Tensor | None = None) -> torch.Tensor:
        """Quantity ``name`` of step ``t``: ``[N, w]`` (``[N]`` for scalars). ``data``: another ``[steps, N * width]`` tensor
        of the same layout (e.g. one rank's part of ``gathered``: ``gathered[:, r]``)."""
        o, w = self._planes[name]
        slab = (self.data if data is None else data)[t]
        v = slab[o:o + self.N * w]
        return v.view(self.N, w) if w > 1 else v

    def done_bytes(self, t: int) -> tuple[torch.Tensor, torch.Tensor]:
        """terminated / truncated of step ``t`` as uint8 ``[N]`` views into the slab's done plane."""
        raw = self.plane("done", t).view(torch.uint8)
        return raw[: self.N], raw[self.N: 2 * self.N]

    def bind(self, buffers, t: int) -> None:
        """Point the outputs of ``buffers``' step launches at slab ``t`` (call before issuing - or capturing - the step)."""
        term, trunc = self.done_bytes(t)
        obs = [self.plane("obs_policy", t) if self.spec.obs[0].dim > 0 else None,
               self.plane("obs_critic", t) if self.spec.obs[1].dim > 0 else None]
        buffers.rebind_outputs(obs=obs, reward=self.plane("reward", t), terminated=term, truncated=trunc)

    # ---- hand-off ---------------------------------------------------------------------------------------------------
    def gather_step(self, t: int, after: "torch.cuda.Stream | None" = None) -> None:
        """Step ``t`` is complete: when it closes a chunk of ``gather_every`` steps, all-gather that chunk on the side stream,
        ordered after everything ``after`` (default: the current stream) has been given so far. Returns immediately; the
        next step can be issued right away."""
        if self.world == 1 or (t + 1) % self.every:
            return
        c = t // self.every
        src_rows = self.data[c * self.every:(c + 1) * self.every].view(-1)
        if self._side is None:   # CPU process groups (gloo tests): synchronous
            dist.all_gather_into_tensor(self._gath[c].view(-1), src_rows, group=self.group)
            return
        src = after if after is not None else torch.cuda.current_stream(self.data.device)
        self._events[t].record(src)
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._events[t])
            dist.all_gather_into_tensor(self._gath[c].view(-1), src_rows, group=self.group)

    def gather_all(self) -> torch.Tensor | None:
        """The un-streamed form: ONE all-gather of the whole rollout on the current stream, into ``[world, steps * N *
        width]`` (the baseline ``gather_step`` is measured against)."""
        if self.world == 1:
            return None
        if self._whole is None:
            self._whole = torch.empty(self.world, self.steps * self.N * self.width, device=self.data.device)
        dist.all_gather_into_tensor(self._whole.view(-1), self.data.view(-1), group=self.group)
        return self._whole

    def finish(self) -> None:
        """Make the current stream wait for every gather issued so far."""
        if self.world > 1 and self._side is not None:
            torch.cuda.current_stream(self.data.device).wait_stream(self._side)

    def global_plane(self, name: str, t: int) -> torch.Tensor:
        """``[world * N, w]`` view-copy of quantity ``name`` of step ``t`` over all ranks (rank-major = global env ids)."""
        if self.world == 1:
            return self.plane(name, t)
        g = self.gathered
        return torch.cat([self.plane(name, t, g[:, r]) for r in range(self.world)], dim=0)


class RolloutBuffer:
    """``[steps, N_local, width]`` fp32 rows filled with ``add`` - the copy-based form (kept for row-major consumers and
    as the baseline of ``RolloutStorage``; nine slice copies per step)."""

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

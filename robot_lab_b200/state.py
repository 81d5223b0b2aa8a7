"""Device-resident state of the MDP step, in the HBM layout the kernels are fastest on.

Layout: small per-env fields are SoA ``[C, N]`` (env innermost -> coalesced sector
loads when a CTA gathers its tile), the sensor streams are AoS rows (contact-force history ``[N, T*B*3]``, ray
hits ``[N, R]``) so that a CTA's tile is one contiguous span for a TMA bulk copy, and policy-facing tensors
(actions, observation rows) are AoS ``[N, D]`` because the policy network consumes rows.

All per-step *inputs* (what physics + sensors + the policy produce) are carved out of one contiguous device
arena, all per-step *results* (observation rows, reward, done masks) out of another, each field 256-byte
aligned: a host-side producer / consumer moves a whole step with ONE copy in each direction.

``layout="aos"`` stores every field IsaacLab-style (``[N, C]``) instead - same kernels through the strided
``RlField`` views; used by tests to cover the generic path and by integrations that hand over PhysX tensors.
"""

from __future__ import annotations

import torch

from . import _native as nat
from .spec import StepSpec

_F32, _U8, _I32 = torch.float32, torch.uint8, torch.int32

# name -> (components as a function of the spec, dtype, preferred layout)
INPUT_FIELDS = {
    "root_pos_w": (lambda s: 3, _F32, "soa"), "root_quat_w": (lambda s: 4, _F32, "soa"),
    "root_lin_vel_w": (lambda s: 3, _F32, "soa"), "root_ang_vel_w": (lambda s: 3, _F32, "soa"),
    "joint_pos": (lambda s: s.J, _F32, "soa"), "joint_vel": (lambda s: s.J, _F32, "soa"),
    "joint_acc": (lambda s: s.J, _F32, "soa"), "applied_torque": (lambda s: s.J, _F32, "soa"),
    "current_air_time": (lambda s: s.Bt, _F32, "soa"), "last_air_time": (lambda s: s.Bt, _F32, "soa"),
    "current_contact_time": (lambda s: s.Bt, _F32, "soa"), "last_contact_time": (lambda s: s.Bt, _F32, "soa"),
    "body_pos_w": (lambda s: s.Ba * 3, _F32, "soa"), "body_lin_vel_w": (lambda s: s.Ba * 3, _F32, "soa"),
    "ray_sensor_pos_z": (lambda s: 1, _F32, "soa"),
    "net_forces_w_history": (lambda s: s.T * s.B * 3, _F32, "aos"),
    "ray_hits_z": (lambda s: s.R, _F32, "aos"),
    "new_action": (lambda s: s.A, _F32, "aos"),
}
MDP_FIELDS = {
    "command": (lambda s: 3, _F32, "soa"), "heading_target": (lambda s: 1, _F32, "soa"),
    "time_left": (lambda s: 1, _F32, "soa"), "is_heading_env": (lambda s: 1, _U8, "soa"),
    "is_standing_env": (lambda s: 1, _U8, "soa"), "metric_error_vel_xy": (lambda s: 1, _F32, "soa"),
    "metric_error_vel_yaw": (lambda s: 1, _F32, "soa"), "episode_length": (lambda s: 1, _I32, "soa"),
    "episode_sums": (lambda s: s.K, _F32, "soa"),
    # manager-internal copies of the action (the policy-facing row tensor is new_action): SoA like every small field, so
    # that the step kernels stage them with one tensor-map copy and write them back as coalesced rows
    "action": (lambda s: s.A, _F32, "soa"), "prev_action": (lambda s: s.A, _F32, "soa"),
    "joint_target": (lambda s: s.J, _F32, "soa"), "joint_vel_target": (lambda s: s.J, _F32, "soa"),
    "step_reward": (lambda s: s.K, _F32, "soa"),
}


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class _Arena:
    """Named typed views into one contiguous uint8 device buffer."""

    def __init__(self, entries: list[tuple[str, tuple[int, ...], torch.dtype]], device: torch.device):
        offs, total = {}, 0
        for name, shape, dtype in entries:
            nbytes = int(torch.tensor([], dtype=dtype).element_size())
            for d in shape:
                nbytes *= d
            offs[name] = (total, shape, dtype, nbytes)
            total = _align(total + nbytes)
        self.nbytes = total
        self.buf = torch.zeros(max(total, 256), dtype=torch.uint8, device=device)
        self.views: dict[str, torch.Tensor] = {}
        for name, (off, shape, dtype, nbytes) in offs.items():
            self.views[name] = self.buf[off:off + nbytes].view(dtype).view(shape)
        self.layout = offs


class StateBuffers:
    """All device tensors one env shard needs: inputs, manager state, outputs."""

    def __init__(self, spec: StepSpec, num_envs: int, device: torch.device | str = "cuda:0", layout: str = "soa"):
        if layout not in ("soa", "aos"):
            raise ValueError(layout)
        self.spec, self.N, self.layout = spec, int(num_envs), layout
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise nat.NativeError("StateBuffers live on a CUDA device - the MDP step has no CPU path")
        N, dev = self.N, self.device
        self.kind: dict[str, str] = {}
        self._cache: dict = {}

        def shape_of(name, comps, pref):
            c = comps(spec)
            kind = pref if layout == "soa" else "aos"
            self.kind[name] = kind
            if c == 1:
                return (N,)
            return (c, N) if kind == "soa" else (N, c)

        self.inputs = _Arena([(n, shape_of(n, c, k), dt) for n, (c, dt, k) in INPUT_FIELDS.items()], dev)
        self.mdp = _Arena([(n, shape_of(n, c, k), dt) for n, (c, dt, k) in MDP_FIELDS.items()], dev)
        out_entries = [(f"obs{g}", (N, grp.dim), _F32) for g, grp in enumerate(spec.obs)]
        out_entries += [("reward", (N,), _F32), ("terminated", (N,), _U8), ("truncated", (N,), _U8)]
        self.outputs = _Arena(out_entries, dev)
        self.t: dict[str, torch.Tensor] = {**self.inputs.views, **self.mdp.views}
        self.obs = [self.outputs.views[f"obs{g}"] if grp.dim > 0 else None for g, grp in enumerate(spec.obs)]
        self.reward = self.outputs.views["reward"]
        self.terminated = self.outputs.views["terminated"]
        self.truncated = self.outputs.views["truncated"]
        self.done_bits = torch.zeros(N, dtype=_U8, device=dev)
        self.reset_ids = torch.zeros(N, dtype=_I32, device=dev)
        self.n_reset = torch.zeros(1, dtype=_I32, device=dev)
        self.step_counter = torch.zeros(1, dtype=torch.int64, device=dev)  # common_step_counter [IL]
        # optional random inputs (noise-as-input mode)
        self.cmd_uniforms: torch.Tensor | None = None
        self.obs_uniforms: list[torch.Tensor | None] = [None, None]
        # reset logging
        # Reset logging: a ring of packed buffers [episode_sum_mean K | done_term_count 8 | metric_mean 2]. The step writes
        # into the current slot; advancing the slot per env step keeps a step's log readable for LOG_RING more steps
        # without any copy (rsl_rl reads the logs of a whole rollout at the end of the iteration).
        kk = max(spec.K, 1)
        self._log_ring = torch.zeros(self.LOG_RING, kk + nat.RL_MAX_DONE_TERMS + 2, device=dev)
        self._log_slot = 0
        self._bind_log_slot()

    LOG_RING = 64

    def _bind_log_slot(self) -> None:
        if getattr(self, "_log_views", None) is None:   # every slot's three views, made once (a view costs ~1.5 us of host time)
            kk = max(self.spec.K, 1)
            self._log_views = [(r, r[:kk], r[kk:kk + nat.RL_MAX_DONE_TERMS], r[kk + nat.RL_MAX_DONE_TERMS:]) for r in self._log_ring]
        self.log_all, self.log_episode_sum_mean, self.log_done_term_count, self.log_metric_mean = self._log_views[self._log_slot]

    def advance_log_slot(self) -> torch.Tensor:
        """Move to the next logging buffer of the ring (call once per env step, before the post-reset launch); returns
        the buffer the coming launch will fill."""
        self._log_slot = (self._log_slot + 1) % self.LOG_RING
        self._bind_log_slot()
        return self.log_all

    # ---- logical <-> device --------------------------------------------------------------------
    def _to_device(self, name: str, logical: torch.Tensor) -> None:
        dst = self.t[name]
        src = logical.to(self.device)
        if src.dtype == torch.bool:
            src = src.to(torch.uint8)
        src = src.to(dst.dtype).reshape(self.N, -1)
        if dst.numel() == 0:
            return
        if dst.dim() == 1:
            dst.copy_(src[:, 0])
        elif self.kind[name] == "soa":
            dst.copy_(src.t())
        else:
            dst.copy_(src)

    def load_logical(self, st: dict) -> None:
        """Copy logical ([N, ...]) tensors (see ``synthetic.make_state``) into the device layout."""
        for name in self.t:
            if name in st:
                self._to_device(name, st[name])
        if st.get("cmd_uniforms") is not None:
            self.cmd_uniforms = st["cmd_uniforms"].to(self.device, torch.float32).contiguous()
        for g, key in enumerate(("obs_uniforms_policy", "obs_uniforms_critic")):
            if st.get(key) is not None and st[key].shape[1] > 0:
                self.obs_uniforms[g] = st[key].to(self.device, torch.float32).contiguous()

    def logical(self, name: str) -> torch.Tensor:
        """One field back in its logical [N, C] / [N] shape (a device tensor; may be a view)."""
        x = self.t[name]
        if x.dim() == 2 and self.kind[name] == "soa":
            x = x.t()
        if name in ("is_heading_env", "is_standing_env"):
            return x.bool()
        if name == "net_forces_w_history":
            return x.reshape(self.N, self.spec.T, self.spec.B, 3)
        if name in ("body_pos_w", "body_lin_vel_w"):
            return x.reshape(self.N, self.spec.Ba, 3)
        return x

    # ---- ctypes views ---------------------------------------------------------------------------------
    def field(self, name: str) -> nat.RlField:
        hit = self._cache.get(("field", name))
        if hit is None:
            hit = self._field(name)
            self._cache[("field", name)] = hit
        return hit

    def _field(self, name: str) -> nat.RlField:
        x = self.t[name]
        if x.numel() == 0:
            return nat.RlField(None, 0, 0)
        if x.dim() == 1:
            return nat.field_of(x, None)
        return nat.field_of(x, self.kind[name])

    # The three pointer structs of a call are built once per StateBuffers: the tensors they point to never move (every
    # update is an in-place copy). Callers that need a private, modifiable struct pass fresh=True.
    def state_view(self, fresh: bool = False) -> nat.RlStateView:
        if not fresh and self._cache.get("state") is not None:
            return self._cache["state"]
        v = nat.RlStateView()
        for name in nat._STATE_FIELDS:
            setattr(v, name, self.field(name))
        if not fresh:
            self._cache["state"] = v
        return v

    def mdp_state(self, fresh: bool = False) -> nat.RlMdpState:
        if not fresh and self._cache.get("mdp") is not None:
            return self._cache["mdp"]
        m = nat.RlMdpState()
        for name in nat._MDP_FIELDS:
            setattr(m, name, self.field(name))
        if not fresh:
            self._cache["mdp"] = m
        return m

    def step_out(self, fresh: bool = False) -> nat.RlStepOut:
        key = ("out", self._log_slot)
        if not fresh and self._cache.get(key) is not None:
            return self._cache[key]
        o = self._step_out()
        if not fresh:
            self._cache[key] = o
        return o

    def rebind_outputs(self, obs: list | None = None, reward: torch.Tensor | None = None,
                       terminated: torch.Tensor | None = None, truncated: torch.Tensor | None = None) -> None:
        """Point the step's results at other device tensors (e.g. the current row of a rollout buffer: the kernels
        write observation rows with any pitch, so a rollout needs no copy of the step outputs)."""
        if obs is not None:
            self.obs = list(obs)
        if reward is not None:
            self.reward = reward
        if terminated is not None:
            self.terminated = terminated
        if truncated is not None:
            self.truncated = truncated
        for k in [k for k in self._cache if isinstance(k, tuple) and k[0] == "out"]:
            del self._cache[k]

    def _step_out(self) -> nat.RlStepOut:
        o = nat.RlStepOut()
        for g in range(nat.RL_NUM_OBS_GROUPS):
            if self.obs[g] is not None:
                o.obs[g] = self.obs[g].data_ptr()
                o.obs_pitch[g] = self.obs[g].stride(0)
        o.reward = self.reward.data_ptr()
        o.terminated = self.terminated.data_ptr()
        o.truncated = self.truncated.data_ptr()
        o.done_bits = self.done_bits.data_ptr()
        o.step_reward = self.field("step_reward")
        o.reset_ids = self.reset_ids.data_ptr()
        o.n_reset = self.n_reset.data_ptr()
        o.reset_log = self.reset_log()
        return o

    def random(self, seed: int = 0, step: int = 0, env_id_offset: int = 0, use_inputs: bool = True,
               use_step_counter: bool = False) -> nat.RlRandom:
        key = ("rnd", seed, step, env_id_offset, use_inputs, use_step_counter, id(self.cmd_uniforms),
               id(self.obs_uniforms[0]), id(self.obs_uniforms[1]))
        hit = self._cache.get(key)
        if hit is not None:
            return hit
        r = self._random(seed, step, env_id_offset, use_inputs, use_step_counter)
        if len(self._cache) > 64:   # rolling seeds / steps: do not grow without bound
            for k in [k for k in self._cache if isinstance(k, tuple) and k[0] == "rnd"]:
                del self._cache[k]
        self._cache[key] = r
        return r

    def _random(self, seed, step, env_id_offset, use_inputs, use_step_counter) -> nat.RlRandom:
        r = nat.RlRandom()
        r.seed, r.step, r.env_id_offset = seed, step, env_id_offset
        if use_step_counter:
            r.step_counter = self.step_counter.data_ptr()
        if use_inputs:
            if self.cmd_uniforms is not None:
                r.cmd_uniforms = self.cmd_uniforms.data_ptr()
            for g in range(nat.RL_NUM_OBS_GROUPS):
                if self.obs_uniforms[g] is not None:
                    r.obs_uniforms[g] = self.obs_uniforms[g].data_ptr()
        return r

    def reset_log(self) -> nat.RlResetLog:
        lg = nat.RlResetLog()
        lg.episode_sum_mean = self.log_episode_sum_mean.data_ptr()
        lg.done_term_count = self.log_done_term_count.data_ptr()
        lg.metric_mean = self.log_metric_mean.data_ptr()
        return lg

    def input_bytes(self) -> int:
        """Payload bytes of the per-step inputs (arena padding excluded)."""
        return sum(v.numel() * v.element_size() for v in self.inputs.views.values())

    def output_bytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.outputs.views.values())

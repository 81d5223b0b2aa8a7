"""Device-resident state of the MDP step, in the HBM layout the kernels are fastest on.

Layout: small per-env fields are SoA ``[C, N]`` (env innermost -> coalesced sector
loads when a CTA gathers its tile), the sensor streams are AoS rows (contact-force history ``[N, T*B*3]``, ray
hits ``[N, R]``) so that a CTA's tile is one contiguous span for a TMA bulk copy, and policy-facing tensors
(actions, observation rows) are AoS ``[N, D]`` because the policy network consumes rows.

``layout="aos"`` stores every field IsaacLab-style (``[N, C]``) instead - same kernels through the strided
``RlField`` views; used by tests to cover the generic path and by integrations that hand over PhysX tensors.
"""

from __future__ import annotations

import torch

from . import _native as nat
from .spec import StepSpec

# name -> (components fn, dtype, group); components as a function of the spec
_SOA_STATE = {
    "root_pos_w": (lambda s: 3, torch.float32), "root_quat_w": (lambda s: 4, torch.float32),
    "root_lin_vel_w": (lambda s: 3, torch.float32), "root_ang_vel_w": (lambda s: 3, torch.float32),
    "joint_pos": (lambda s: s.J, torch.float32), "joint_vel": (lambda s: s.J, torch.float32),
    "joint_acc": (lambda s: s.J, torch.float32), "applied_torque": (lambda s: s.J, torch.float32),
    "current_air_time": (lambda s: s.Bt, torch.float32), "last_air_time": (lambda s: s.Bt, torch.float32),
    "current_contact_time": (lambda s: s.Bt, torch.float32), "last_contact_time": (lambda s: s.Bt, torch.float32),
    "body_pos_w": (lambda s: s.Ba * 3, torch.float32), "body_lin_vel_w": (lambda s: s.Ba * 3, torch.float32),
    "ray_sensor_pos_z": (lambda s: 1, torch.float32),
}
_AOS_STATE = {
    "net_forces_w_history": (lambda s: s.T * s.B * 3, torch.float32),
    "ray_hits_z": (lambda s: s.R, torch.float32),
}
_SOA_MDP = {
    "command": (lambda s: 3, torch.float32), "heading_target": (lambda s: 1, torch.float32),
    "time_left": (lambda s: 1, torch.float32), "is_heading_env": (lambda s: 1, torch.uint8),
    "is_standing_env": (lambda s: 1, torch.uint8), "metric_error_vel_xy": (lambda s: 1, torch.float32),
    "metric_error_vel_yaw": (lambda s: 1, torch.float32), "episode_length": (lambda s: 1, torch.int32),
    "episode_sums": (lambda s: s.K, torch.float32),
}
_AOS_MDP = {"action": (lambda s: s.A, torch.float32), "prev_action": (lambda s: s.A, torch.float32)}


class StateBuffers:
    """All device tensors one env shard needs: inputs, manager state, outputs."""

    def __init__(self, spec: StepSpec, num_envs: int, device: torch.device | str = "cuda:0", layout: str = "soa"):
        if layout not in ("soa", "aos"):
            raise ValueError(layout)
        self.spec, self.N, self.layout = spec, int(num_envs), layout
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise nat.NativeError("StateBuffers live on a CUDA device - the MDP step has no CPU path")
        N = self.N
        self.t: dict[str, torch.Tensor] = {}
        self.kind: dict[str, str] = {}

        def alloc(name, comps, dtype, kind):
            c = comps(spec)
            if kind == "soa" and layout == "soa":
                self.t[name] = torch.zeros((c, N) if c != 1 else (N,), dtype=dtype, device=self.device)
                self.kind[name] = "soa"
            else:
                self.t[name] = torch.zeros((N, c) if c != 1 else (N,), dtype=dtype, device=self.device)
                self.kind[name] = "aos"

        for name, (comps, dt) in _SOA_STATE.items():
            alloc(name, comps, dt, "soa")
        for name, (comps, dt) in _AOS_STATE.items():
            alloc(name, comps, dt, "aos")
        for name, (comps, dt) in _SOA_MDP.items():
            alloc(name, comps, dt, "soa")
        for name, (comps, dt) in _AOS_MDP.items():
            alloc(name, comps, dt, "aos")
        alloc("new_action", lambda s: s.A, torch.float32, "aos")
        alloc("joint_target", lambda s: s.J, torch.float32, "soa")
        alloc("step_reward", lambda s: s.K, torch.float32, "soa")
        # outputs
        dev = self.device
        self.obs = [torch.zeros(N, max(g.dim, 1), device=dev)[:, : g.dim].contiguous() if g.dim > 0 else None
                    for g in spec.obs]
        self.reward = torch.zeros(N, device=dev)
        self.terminated = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.done_bits = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.reset_ids = torch.zeros(N, dtype=torch.int32, device=dev)
        self.n_reset = torch.zeros(1, dtype=torch.int32, device=dev)
        # optional random inputs (noise-as-input mode)
        self.cmd_uniforms: torch.Tensor | None = None
        self.obs_uniforms: list[torch.Tensor | None] = [None, None]
        # reset logging
        self.log_episode_sum_mean = torch.zeros(max(spec.K, 1), device=dev)
        self.log_done_term_count = torch.zeros(nat.RL_MAX_DONE_TERMS, device=dev)
        self.log_metric_mean = torch.zeros(2, device=dev)

    # ---- logical <-> device --------------------------------------------------------------------
    def _to_device(self, name: str, logical: torch.Tensor) -> None:
        dst = self.t[name]
        src = logical.to(self.device)
        if src.dtype == torch.bool:
            src = src.to(torch.uint8)
        src = src.to(dst.dtype).reshape(self.N, -1)
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
        """One field back in its logical [N, C] / [N] shape (a device tensor; may be a copy)."""
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
        x = self.t[name]
        if x.numel() == 0:
            return nat.RlField(None, 0, 0)
        if x.dim() == 1:
            return nat.field_of(x, None)
        return nat.field_of(x, self.kind[name])

    def state_view(self) -> nat.RlStateView:
        v = nat.RlStateView()
        for name in nat._STATE_FIELDS:
            setattr(v, name, self.field(name))
        return v

    def mdp_state(self) -> nat.RlMdpState:
        m = nat.RlMdpState()
        for name in nat._MDP_FIELDS:
            setattr(m, name, self.field(name))
        return m

    def step_out(self) -> nat.RlStepOut:
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
        return o

    def random(self, seed: int = 0, step: int = 0, env_id_offset: int = 0, use_inputs: bool = True) -> nat.RlRandom:
        r = nat.RlRandom()
        r.seed, r.step, r.env_id_offset = seed, step, env_id_offset
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
        """Bytes of the per-step inputs (physics/sensor state + new action) - what an e2e step copies H2D."""
        names = list(nat._STATE_FIELDS) + ["new_action"]
        return sum(self.t[n].numel() * self.t[n].element_size() for n in names)

    def output_bytes(self) -> int:
        outs = [o for o in self.obs if o is not None] + [self.reward, self.terminated, self.truncated]
        return sum(o.numel() * o.element_size() for o in outs)

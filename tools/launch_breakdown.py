"""Steady-state cost of each launch of one env step, measured back to back with CUDA events (production kernels,
no debug buffer): process_action | rl_step(ALL|SKIP) | rl_step(RESET|COMMAND|OBS on reset ids) and a few phase
subsets of the fused kernel. Rotates over independent state sets larger than L2.

Usage (GPU box): python tools/launch_breakdown.py [num_envs] [warps] [task_key] [envs_per_cta: 32] [--short] [--pdl]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))

import torch  # noqa: E402

import helpers as H  # noqa: E402
from robot_lab_b200 import _native as nat  # noqa: E402
from robot_lab_b200.engine import MdpStepEngine  # noqa: E402
from robot_lab_b200.synthetic import make_state  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W = int(sys.argv[2]) if len(sys.argv) > 2 else 16
key = sys.argv[3] if len(sys.argv) > 3 else "go2_rough"
EPC = int(sys.argv[4]) if len(sys.argv) > 4 else 0
SHORT = "--short" in sys.argv   # only the three launches of an env step and their sum
cfg, spec = H.make_spec(key)
eng = MdpStepEngine(spec, "cuda:0")
eng.set_launch_config(W, EPC)
n_sets = max(2, min(24, int(400e6 / (N * 3500)) + 1))
sets = []
for i in range(n_sets):
    b = eng.new_buffers(N)
    b.load_logical(make_state(spec, N, seed=1234 + i))
    b.cmd_uniforms, b.obs_uniforms = None, [None, None]
    sets.append(b)
ALL = nat.PHASE_ALL | nat.PHASE_SKIP_DONE_ENVS
POST = nat.PHASE_RESET | nat.PHASE_COMMAND | nat.PHASE_OBS


def run_step(b):
    eng.step(b, phases=ALL, use_random_inputs=False, use_step_counter=True)


def run_post(b):
    eng.step(b, phases=POST, use_random_inputs=False, use_step_counter=True, env_ids=b.reset_ids, n_env_ids=b.n_reset)


net_forces = torch.randn(N, spec.B, 3, device="cuda:0") * (torch.rand(N, spec.B, 1, device="cuda:0") < 0.3)
from robot_lab_b200.cfg import ResetStateCfg  # noqa: E402

reset_cfg = ResetStateCfg.go2_rough()
origins = torch.zeros(N, 3, device="cuda:0")
CASES = {
    "reset_scene_state (root + joints of the done envs)": lambda b: eng.reset_scene_state(b, reset_cfg, origins, use_step_counter=True),
    "contact_sensor_update (history roll + timers)": lambda b: eng.contact_sensor_update(b, net_forces, 0.005),
    "contact_sensor_update (ring slot + timers)": lambda b: eng.contact_sensor_update(b, net_forces, 0.005, ring_slot=0),
    "process_action": lambda b: eng.process_action(b),
    "rl_step DONES|REWARDS|COMPACT (pre-reset)": lambda b: eng.step_pre_reset(b, use_random_inputs=False, use_step_counter=True),
    "rl_step RESET|COMMAND|OBS, all envs (post-reset)": lambda b: eng.step_post_reset(b, use_random_inputs=False, use_step_counter=True),
    "env step: process_action + pre + post": lambda b: (eng.process_action(b), eng.step_pre_reset(b, use_random_inputs=False, use_step_counter=True),
                                                       eng.step_post_reset(b, use_random_inputs=False, use_step_counter=True)),
    "rl_step ALL|SKIP_DONE": run_step,
    "rl_step RESET|COMMAND|OBS (reset ids)": run_post,
    "rl_step DONES only": lambda b: eng.step(b, phases=nat.PHASE_DONES, use_random_inputs=False),
    "rl_step DONES|REWARDS": lambda b: eng.step(b, phases=nat.PHASE_DONES | nat.PHASE_REWARDS, use_random_inputs=False),
    "rl_step OBS only": lambda b: eng.step(b, phases=nat.PHASE_OBS, use_random_inputs=False),
    "rl_step COMMAND only": lambda b: eng.step(b, phases=nat.PHASE_COMMAND, use_random_inputs=False),
    "env step, older form: process_action + ALL|SKIP + reset ids": lambda b: (eng.process_action(b), run_step(b), run_post(b)),
}
if SHORT:
    CASES = {k: v for k, v in CASES.items() if k.startswith(("process_action", "rl_step DONES|REWARDS|COMPACT", "rl_step RESET|COMMAND|OBS, all", "env step: "))}
for b in sets:  # reset ids valid for the post-reset case
    run_step(b)
torch.cuda.synchronize()
print(f"{key} N={N} warps={W} envs_per_cta={EPC or 32} state sets={n_sets}; mean reset envs per step: "
      f"{sum(int(b.n_reset.item()) for b in sets) / len(sets):.1f}")
for pdl in ((False, True) if "--pdl" in sys.argv else (False,)):
    eng.set_pdl(pdl)
    for name, fn in CASES.items():
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for b in sets:
                fn(b)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for b in sets:
                    fn(b)
            for _ in range(3):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record(s)
            for _ in range(reps):
                g.replay()
            e1.record(s)
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * len(sets))
        print(f"  pdl={int(pdl)}  {name:58s} {us:8.2f} us")

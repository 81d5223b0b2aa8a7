"""A/B of one tile per CTA (512 threads, 32 envs) against two tiles per CTA (1024 threads, 64 envs) on the same box,
same state sets, alternating: the three launches of an env step and their sum, CUDA graphs of back-to-back launches
rotating over state sets larger than L2, CUDA events (the method of tools/launch_breakdown.py).

Usage (GPU box): python tools/two_tile_ab.py [task_key] [num_envs ...]     (default: go2_rough 4096 16384)
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))

import torch  # noqa: E402

import helpers as H  # noqa: E402
from robot_lab_b200.engine import MdpStepEngine  # noqa: E402
from robot_lab_b200.synthetic import make_state  # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "go2_rough"
sizes = [int(x) for x in sys.argv[2:]] or [4096, 16384]
cfg, spec = H.make_spec(key)
eng = MdpStepEngine(spec, "cuda:0")
kw = dict(use_random_inputs=False, use_step_counter=True)
CASES = {
    "process_action": lambda b: eng.process_action(b),
    "pre_reset (DONES|REWARDS|COMPACT)": lambda b: eng.step_pre_reset(b, **kw),
    "post_reset (RESET|COMMAND|OBS)": lambda b: eng.step_post_reset(b, **kw),
    "env_step": lambda b: (eng.process_action(b), eng.step_pre_reset(b, **kw), eng.step_post_reset(b, **kw)),
}


def measure(sets, fn, reps=20):
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
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(sets))


result = {"task": key, "device": torch.cuda.get_device_name(0), "runs": []}
for N in sizes:
    n_sets = max(2, min(24, int(400e6 / (N * 3500)) + 1))
    sets = []
    for i in range(n_sets):
        b = eng.new_buffers(N)
        b.load_logical(make_state(spec, N, seed=1234 + i))
        b.cmd_uniforms, b.obs_uniforms = None, [None, None]
        sets.append(b)
    for epc in (32, 64, 32, 64):
        eng.set_launch_config(16, epc)
        row = {"num_envs": N, "state_sets": n_sets, **eng.launch_config()}
        for name, fn in CASES.items():
            row[name + "_us"] = round(measure(sets, fn), 3)
        result["runs"].append(row)
        print(json.dumps(row), flush=True)
    del sets
    torch.cuda.empty_cache()
print("RESULT " + json.dumps(result))

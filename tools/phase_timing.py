"""Per-CTA phase timeline of the fused step kernel (clock64 stamps, see rl_ctx_set_debug_buffer).

Usage (GPU box): python tools/phase_timing.py [num_envs] [warps] [task_key]
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
GW = sys.argv[2] if len(sys.argv) > 2 else "16x16"
G_, W = (int(x) for x in GW.split("x"))
key = sys.argv[3] if len(sys.argv) > 3 else "go2_rough"
PHASES = {"all": nat.PHASE_ALL | nat.PHASE_SKIP_DONE_ENVS, "dones": nat.PHASE_DONES, "rewards": nat.PHASE_REWARDS,
          "obs": nat.PHASE_OBS, "command": nat.PHASE_COMMAND, "dones+compact": nat.PHASE_DONES | nat.PHASE_COMPACT,
          "dones+rewards": nat.PHASE_DONES | nat.PHASE_REWARDS, "empty": 0x8000, "ctx": 0x4000}
phase_name = sys.argv[4] if len(sys.argv) > 4 else "all"
cfg, spec = H.make_spec(key)
eng = MdpStepEngine(spec, "cuda:0")
eng.set_launch_config(G_, W)
sets = []
for i in range(8):
    b = eng.new_buffers(N)
    b.load_logical(make_state(spec, N, seed=1234 + i))
    b.cmd_uniforms, b.obs_uniforms = None, [None, None]
    sets.append(b)
grid = (((N + 31) // 32 + W - 1) // W) * G_
dbg = torch.zeros(grid, 8, dtype=torch.int64, device="cuda:0")
PH = PHASES[phase_name]
for it in range(6):
    for b in sets:
        eng.step(b, phases=PH, use_random_inputs=False)
torch.cuda.synchronize()
eng.set_debug_buffer(dbg)
acc = torch.zeros(1)
mx = torch.zeros(1)
per_group = torch.zeros(G_)
tot = []
reps = 8
for it in range(reps):
    b = sets[it % len(sets)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng.step(b, phases=PH, use_random_inputs=False)
    ev1.record()
    torch.cuda.synchronize()
    d = dbg.cpu().double()
    if phase_name in ("empty", "ctx"):
        k = 1 if phase_name == "empty" else 2
        print("   start->stamp%d cycles: mean %.0f max %.0f" % (k, (d[:, k] - d[:, 0]).mean().item(), (d[:, k] - d[:, 0]).max().item()),
              " stamp1: mean %.0f" % (d[:, 1] - d[:, 0]).mean().item(), " event us %.1f" % (ev0.elapsed_time(ev1) * 1e3))
        continue
    span = (d[:, 3] - d[:, 0])            # start -> this CTA's tasks done
    per_group += span.view(-1, G_).mean(0).float()
    t0 = d[:, 0].min()
    tot.append((span.mean().item(), span.max().item(), (d[:, 3].max() - t0).item(), ev0.elapsed_time(ev1) * 1e3))
print(f"{key} N={N} groups x warps={G_}x{W} grid={grid} phases={phase_name}")
if phase_name in ("empty", "ctx"):
    sys.exit(0)
print("  task cycles per group (mean over tile blocks):", [int(x) for x in (per_group / reps).tolist()])
print("  per-CTA task cycles (mean, max), first start -> last tasks done, event us:", [tuple(round(x, 1) for x in t) for t in tot[-3:]])

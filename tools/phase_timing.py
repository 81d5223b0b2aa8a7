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
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
key = sys.argv[3] if len(sys.argv) > 3 else "go2_rough"
PHASES = {"all": nat.PHASE_ALL | nat.PHASE_SKIP_DONE_ENVS, "dones": nat.PHASE_DONES, "rewards": nat.PHASE_REWARDS,
          "obs": nat.PHASE_OBS, "command": nat.PHASE_COMMAND, "dones+compact": nat.PHASE_DONES | nat.PHASE_COMPACT,
          "dones+rewards": nat.PHASE_DONES | nat.PHASE_REWARDS,
          "post": nat.PHASE_RESET | nat.PHASE_COMMAND | nat.PHASE_OBS}
phase_name = sys.argv[4] if len(sys.argv) > 4 else "all"
cfg, spec = H.make_spec(key)
eng = MdpStepEngine(spec, "cuda:0")
eng.set_launch_config(W)
sets = []
for i in range(8):
    b = eng.new_buffers(N)
    b.load_logical(make_state(spec, N, seed=1234 + i))
    b.cmd_uniforms, b.obs_uniforms = None, [None, None]
    sets.append(b)
grid = (N + 31) // 32
dbg = torch.zeros(grid, nat.RL_DEBUG_STRIDE, dtype=torch.int64, device="cuda:0")
PH = PHASES[phase_name]
POST = phase_name == "post"


def launch(b):
    if POST:
        eng.step(b, phases=PH, use_random_inputs=False, env_ids=b.reset_ids, n_env_ids=b.n_reset)
    else:
        eng.step(b, phases=PH, use_random_inputs=False)


if POST:
    for b in sets:
        eng.step(b, phases=nat.PHASE_ALL | nat.PHASE_SKIP_DONE_ENVS, use_random_inputs=False)
    torch.cuda.synchronize()
    print("reset envs per set:", [int(b.n_reset.item()) for b in sets])
for it in range(6):
    for b in sets:
        launch(b)
torch.cuda.synchronize()
eng.set_debug_buffer(dbg)
names = ["start->loads issued", "loads issued->tile resident", "tile resident->stage1 done", "stage1->stage2 done",
         "stage2->stores issued", "stores->compaction", "compaction->exit"]
acc = torch.zeros(7)
mx = torch.zeros(7)
sched = eng.schedule()
task_acc = torch.zeros(len(sched))
enter_acc = torch.zeros(W)
sub_acc = torch.zeros(5)
tot = []
reps = 8
for it in range(reps):
    b = sets[it % len(sets)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    launch(b)
    ev1.record()
    torch.cuda.synchronize()
    dall = dbg.cpu().double()
    if POST:
        dall = dall[: (int(b.n_reset.item()) + 31) // 32]
    d = dall[:, :8]
    task_acc += dall[:, 8:8 + len(sched)].mean(0).float()
    enter_acc += (dall[:, 8 + nat.RL_MAX_TASKS:8 + nat.RL_MAX_TASKS + W] - dall[:, 2:3]).mean(0).float()
    sub = dall[:, 8 + nat.RL_MAX_TASKS + 32:8 + nat.RL_MAX_TASKS + 37] - dall[:, 0:1]
    sub_acc += sub.mean(0).float()
    dur = d[:, 1:] - d[:, :-1]
    dur[dur.abs() > 1e8] = 0  # stamps a CTA skipped (early return of the compacting CTA)
    acc += dur.mean(0).float()
    mx = torch.maximum(mx, dur.max(0).values.float())
    span = (d[:, 5] - d[:, 0])
    tot.append((span.mean().item(), span.max().item(), ev0.elapsed_time(ev1) * 1e3))
print(f"{key} N={N} warps={W} grid={grid} phases={phase_name}")
for n, a_, m_ in zip(names, acc / reps, mx):
    print(f"  {n:32s} mean {a_:9.0f} cyc   max {m_:9.0f} cyc")
print("  cycles since kernel start: bulk copies issued %.0f, per-field copies issued %.0f, span loads issued %.0f; "
      "stage 2: late terms finished %.0f, reward summed %.0f" % tuple((sub_acc / reps).tolist()))
print("  per-CTA total cycles (mean, max), event us:", [tuple(round(x, 1) for x in t) for t in tot[-3:]])

if phase_name == "all":
    rnames = [t.name for t in spec.rewards]
    per_warp = {}
    for t, cyc in zip(sched, (task_acc / reps).tolist()):
        if t["kind"] == 0:
            label = f"reward {rnames[t['a']]}" + (f" [bodies {t['lo']}:{t['hi']}]" if t["late"] else "")
        elif t["kind"] == 1:
            label = f"obs g{t['a']} term {spec.obs[t['a']].terms[t['b']].name} cols {t['lo']}:{t['hi']}"
        else:
            label = "terminations" if t["kind"] == 2 else "command (+ its obs columns)"
        per_warp.setdefault(t["owner"], []).append((label, cyc))
    print("  stage-1 tasks per warp (mean cycles); 'enter' = tile resident -> warp starts its first task (ctx set-up)")
    for w in sorted(per_warp):
        tot_w = sum(c for _, c in per_warp[w])
        print(f"   warp {w:2d}: enter {enter_acc[w] / reps:6.0f}  tasks {tot_w:7.0f}  " + "; ".join(f"{l} {c:.0f}" for l, c in per_warp[w]))

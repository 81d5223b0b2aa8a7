"""One cold launch of the two step kernels (Go2 rough, 4096 envs) for each launch configuration - one tile per CTA
(128 x 512 threads), then two tiles per CTA (64 x 1024 threads) - bracketed by cudaProfilerStart/Stop:

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/two_tiles \
      python tools/ncu_two_tiles.py
  ncu -i gpurun_out/two_tiles.ncu-rep --page raw --csv > gpurun_out/two_tiles_raw.csv
  python tools/ncu_two_tiles.py --summarise gpurun_out/two_tiles_raw.csv     (no GPU needed)
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))

KEYS = [
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("gpu__time_duration.sum", "duration us"),
    ("smsp__inst_executed.sum", "warp instructions"), ("sm__inst_executed.avg.per_cycle_active", "IPC per active SM cycle"),
    ("sm__warps_active.avg.per_cycle_active", "warps active per SM cycle"),
    ("dram__bytes_read.sum", "DRAM read (MB)"),
] + [(f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio", f"stall {k}") for k in
     ("barrier", "no_instruction", "short_scoreboard", "long_scoreboard", "wait", "branch_resolving", "mio_throttle",
      "math_pipe_throttle", "membar", "not_selected", "lg_throttle", "dispatch_stall")]


def summarise(path):
    import csv

    rows = list(csv.reader(open(path)))
    hdr, body = rows[0], [r for r in rows[2:] if len(r) == len(rows[0])]
    col = {h: i for i, h in enumerate(hdr)}
    names = [f"launch {r[col['ID']]}" for r in body]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(body))
    for key, label in KEYS:
        if key in col:
            print(f"| {label} | " + " | ".join(r[col[key]] for r in body) + " |")


if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
    summarise(sys.argv[2])
    raise SystemExit(0)

import torch  # noqa: E402

import helpers as H  # noqa: E402
from robot_lab_b200.engine import MdpStepEngine  # noqa: E402
from robot_lab_b200.synthetic import make_state  # noqa: E402

N = 4096
cfg, spec = H.make_spec("go2_rough")
eng = MdpStepEngine(spec, "cuda:0")
sets = []
for i in range(4):
    b = eng.new_buffers(N)
    b.load_logical(make_state(spec, N, seed=1234 + i))
    b.cmd_uniforms, b.obs_uniforms = None, [None, None]
    sets.append(b)
rng = dict(seed=42, use_random_inputs=False, use_step_counter=True)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for epc in (32, 64):
    eng.set_launch_config(16, epc)
    for b in sets[:2]:      # module load, instruction caches
        eng.step_pre_reset(b, **rng)
        eng.step_post_reset(b, **rng)
torch.cuda.synchronize()
for epc, b in ((32, sets[2]), (64, sets[3])):
    eng.set_launch_config(16, epc)
    flush.fill_(1)          # evict the set from L2: the captured launches read HBM
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.step_pre_reset(b, **rng)
    eng.step_post_reset(b, **rng)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("captured pre/post launches for 32 and 64 envs per CTA")

"""Launch every SURVEY.md 8(f) neighbour kernel a few times on synthetic Go2-rough state - the target of the ncu
captures in profiles/ (and a quick CUDA-event timing when run alone).

Usage (GPU box): python tools/neighbour_probe.py [num_envs] [repeats]
  ncu --set full --clock-control none --import-source on -k regex:"actuator|height_scan|terrain|contact_sensor" \
      -c 8 -o gpurun_out/neighbours python tools/neighbour_probe.py 4096 2
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))

import torch  # noqa: E402

import helpers as H  # noqa: E402
from robot_lab_b200 import terrain as terrain_host  # noqa: E402
from robot_lab_b200.cfg import RayCasterCfg, TerrainCfg  # noqa: E402
from robot_lab_b200.engine import MdpStepEngine  # noqa: E402
from robot_lab_b200.synthetic import make_state  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
REP = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cfg, spec = H.make_spec("go2_rough")
eng = MdpStepEngine(spec, "cuda:0")
n_sets = max(2, min(24, int(400e6 / (N * 3500)) + 1))   # rotate over more than the L2 holds
sets = []
for i in range(n_sets):
    b = eng.new_buffers(N)
    b.load_logical(make_state(spec, N, seed=1234 + i))
    sets.append(b)
ter = TerrainCfg()
nx = int(round((ter.num_rows * ter.size[0] + 2 * ter.border_width) / ter.horizontal_scale)) + 1
ny = int(round((ter.num_cols * ter.size[1] + 2 * ter.border_width) / ter.horizontal_scale)) + 1
g = torch.Generator().manual_seed(7)
heights = torch.nn.functional.interpolate((torch.rand(nx // 8 + 2, ny // 8 + 2, generator=g) * 0.8)[None, None], size=(nx, ny),
                                          mode="bilinear", align_corners=True)[0, 0].contiguous().cuda()
hf = terrain_host.HeightFieldBuffers(heights, -0.5 * (nx - 1) * 0.1, -0.5 * (ny - 1) * 0.1, 0.1,
                                     terrain_host.grid_pattern_ray_starts(RayCasterCfg()).cuda())
pit = TerrainCfg(sub_terrains=("pyramid_stairs", "pits", "boxes", "random_rough", "hf_pyramid_slope"),
                 proportions=(0.2, 0.15, 0.25, 0.3, 0.1))
grid = terrain_host.TerrainGridBuffers.create(pit, "pits", "cuda:0")
was = torch.zeros(N, dtype=torch.uint8, device="cuda")
forces = torch.randn(N, spec.B, 3, device="cuda")
KERNELS = {
    "rl_actuator_step": (lambda b: eng.actuator_step(b), 4 * 4 * spec.J),
    "rl_contact_sensor_update(ring)": (lambda b: eng.contact_sensor_update(b, forces, 0.005, ring_slot=1),
                                       4 * (6 * spec.B + 6 * spec.Bt)),
    "rl_height_scan_cast": (lambda b: eng.height_scan_cast(b, hf), 4 * (8 + spec.R)),
    "rl_command_pit_restrict": (lambda b: eng.command_pit_restrict(b, grid, was, seed=1, use_random_inputs=False), 10),
}
for name, (fn, nbytes) in KERNELS.items():
    for b in sets[:2]:
        fn(b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gk = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.graph(gk, stream=s):
        for i in range(n_sets):
            fn(sets[i])
    with torch.cuda.stream(s):
        gk.replay()
        e0.record(s)
        for _ in range(REP):
            gk.replay()
        e1.record(s)
    s.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (REP * n_sets)
    print(f"{name:34s} N={N:6d}  {us:8.2f} us   {nbytes} B/env -> {nbytes * N / us / 1e3:8.1f} GB/s")

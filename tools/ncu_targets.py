"""One launch of every kernel of the library on cold synthetic Go2-rough state (4096 envs), bracketed by
cudaProfilerStart/Stop - the target of the `ncu --set full` captures summarised in profiles/.

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/full \
      python tools/ncu_targets.py [num_envs]
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
cfg, spec = H.make_spec("go2_rough")
eng = MdpStepEngine(spec, "cuda:0")
sets = []
for i in range(3):
    b = eng.new_buffers(N)
    b.load_logical(make_state(spec, N, seed=1234 + i))
    b.cmd_uniforms, b.obs_uniforms = None, [None, None]
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
rng = dict(seed=42, use_random_inputs=False, use_step_counter=True)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def everything(b):
    eng.process_action(b)
    eng.actuator_step(b)
    eng.contact_sensor_update(b, forces, 0.005, ring_slot=1)
    eng.height_scan_cast(b, hf)
    eng.step_pre_reset(b, **rng)
    eng.step_post_reset(b, **rng)
    eng.command_pit_restrict(b, grid, was, seed=1, use_random_inputs=False)


for b in sets[:2]:      # first calls: scratch allocation, module load, instruction caches of the other sets
    everything(b)
flush.fill_(1)          # evict the third set from L2: the captured launches read HBM
torch.cuda.synchronize()
torch.cuda.profiler.start()
everything(sets[2])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("captured one launch of each kernel at", N, "envs")

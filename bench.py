#!/usr/bin/env python
"""Benchmark of the per-step MDP pipeline: env-steps/s on synthetic 4096-env Unitree-Go2 rough-velocity state.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

One "step" = one pass of the hot path over one batch of synthetic env state, in the order of
ManagerBasedRLEnv.step() [IL]: process_action -> step kernel, launch 1 (terminations, rewards, reset-id
compaction) -> step kernel, launch 2 (manager reset + logging means of the done envs, command, observations). Physics / sensors are the *producer* of the state buffers and are not part of this
tier: the state is synthetic (robot_lab_b200.synthetic) and resident in HBM before the timed region.

Timing rules followed: W >= 3 warm-up steps; the step rotates over S >= 16 independent state sets whose
combined footprint exceeds the 126 MB L2 (config.l2_policy says so); CUDA events on the launching stream with
barrier + synchronize on both sides; max over ranks; nvidia-smi clocks sampled during the timed region.
Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

TASK_DEFAULT = "RobotLab-Isaac-Velocity-Rough-Unitree-Go2-v0"
METRIC = "env-steps/sec (4096-env Go2 rough-velocity)"
UNIT = "env-steps/s"
ROLLOUT = 24  # steps per PPO rollout (GO2/agents/rsl_rl_ppo_cfg.py:11) = steps per captured CUDA graph


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2400)
    p.add_argument("--warmup", type=int, default=240)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--task", default=TASK_DEFAULT)
    p.add_argument("--num-envs", type=int, default=4096, help="envs per GPU (weak scaling)")
    p.add_argument("--sets", type=int, default=24, help="independent state sets the step rotates over (L2 defeat)")
    p.add_argument("--warps", type=int, default=0, help="warps per tile (4/8/16; a tile is 32 envs)")
    p.add_argument("--envs-per-cta", type=int, default=0, choices=[0, 32, 64],
                   help="32 = one tile per CTA (default), 64 = two tiles per CTA (1024 threads)")
    p.add_argument("--pdl", action="store_true", help="programmatic dependent launch between the kernels (measured slower)")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the bounded CPU baseline sample")
    p.add_argument("--skip-handoff", action="store_true")
    p.add_argument("--no-neighbours", action="store_true", help="skip the side measurements of the SURVEY 8(f) kernels")
    return p.parse_args()


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle port (eager fp32 torch, the reference's own op stream) on host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_step_fn(spec, st):
    from oracle import mdp_port as port

    rnd = {"cmd_uniforms": st["cmd_uniforms"], "obs_uniforms_policy": st["obs_uniforms_policy"],
           "obs_uniforms_critic": st["obs_uniforms_critic"]}

    def one_step():
        action, prev, _target = port.process_action(spec, st, st["new_action"])
        s1 = dict(st)
        s1["action"], s1["prev_action"] = action, prev
        out = port.step(spec, s1, rnd, skip_done_envs=True)
        s2 = dict(s1)
        s2.update({k: out[k] for k in ("command", "heading_target", "time_left", "is_heading_env", "is_standing_env",
                                       "metric_error_vel_xy", "metric_error_vel_yaw", "episode_length", "episode_sums")})
        s3, _log = port.reset_envs(spec, s2, out["reset_ids"], out["done_bits"], rnd)
        s2.update(s3)
        mask = torch.zeros(st["root_quat_w"].shape[0], dtype=torch.bool)
        mask[out["reset_ids"].long()] = True
        s2.update(port.compute_command(spec, s2, rnd, active=mask))
        port.compute_obs_group(spec, 0, s2, rnd)
        port.compute_obs_group(spec, 1, s2, rnd)
        return out["reward"]

    return one_step


def time_cpu(spec, num_envs: int, steps: int, warmup: int, budget_s: float | None):
    from robot_lab_b200.synthetic import make_state

    st = make_state(spec, num_envs)
    fn = cpu_step_fn(spec, st)
    # eager torch on [4096, C] tensors is dispatch-bound: more intra-op threads can be slower. Give the CPU arm
    # its best case: sweep the thread count (2 steps each) and keep the fastest.
    ncpu = os.cpu_count() or 1
    sweep, best = {}, None
    for nt in sorted({1, 2, 4, 8, 16, 32, 64, ncpu}):
        if nt > ncpu:
            continue
        torch.set_num_threads(nt)
        fn()
        t0 = time.perf_counter()
        fn()
        sweep[nt] = time.perf_counter() - t0
        if best is None or sweep[nt] < sweep[best]:
            best = nt
    cores = best
    torch.set_num_threads(cores)
    for _ in range(max(1, warmup)):
        fn()
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        fn()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": num_envs * done / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} full MDP steps of {num_envs} envs ({dt:.1f} s), eager fp32 torch oracle port, "
                      f"torch.set_num_threads({cores}) = fastest of a sweep over {sorted(sweep)} threads on {ncpu} logical CPUs",
            "ms_per_step": 1e3 * dt / done, "steps": done}


def run_reference(args, spec, rank: int, world: int):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference's own
    Python files cannot travel to the GPU box) on this box's host cores. Rank 0 only."""
    if rank != 0:
        return
    # each "step" = one full MDP step of num_envs envs; bounded so that the run ends within minutes
    steps = min(args.steps, 200)
    res = time_cpu(spec, args.num_envs, steps, min(args.warmup, 3), budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": res["steps"], "warmup": min(args.warmup, 3), "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.task}, {args.num_envs} envs, full MDP step on host CPU"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from robot_lab_b200.spec import compact_layout, compile_step_spec
    from robot_lab_b200.tasks import make_env_cfg

    cfg = make_env_cfg(args.task, num_envs=args.num_envs)
    spec = compile_step_spec(cfg, compact_layout(cfg))

    if args.impl == "reference":
        run_reference(args, spec, rank, world)
        return

    import torch.distributed as dist

    from robot_lab_b200 import _native as nat
    from robot_lab_b200.engine import MdpStepEngine
    from robot_lab_b200.synthetic import make_state

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MDP step has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # NCCL's version banner goes to stdout: keep stdout to the ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    N, S, K, W = args.num_envs, max(1, args.sets), args.steps, max(3, args.warmup)
    eng = MdpStepEngine(spec, dev)
    if args.warps or args.envs_per_cta:
        eng.set_launch_config(args.warps, args.envs_per_cta)
    eng.set_pdl(args.pdl)
    sets = []
    for i in range(S):
        b = eng.new_buffers(N)
        b.load_logical(make_state(spec, N, seed=1234 + 1000 * i, rank=rank))
        b.cmd_uniforms, b.obs_uniforms = None, [None, None]  # production mode: in-kernel Philox, no noise bytes
        sets.append(b)
    env_off = rank * N
    rng = dict(seed=42 + rank, env_id_offset=env_off, use_random_inputs=False, use_step_counter=True)

    def one_step(b):
        # ManagerBasedRLEnv.step() [IL]: process_action -> (physics: the synthetic provider changes nothing) ->
        # terminations, rewards, reset ids -> (external reset) -> manager reset, command, observations
        eng.process_action(b)
        eng.step_pre_reset(b, **rng)
        eng.step_post_reset(b, **rng)

    launches_per_step = 3
    stream = torch.cuda.Stream(device=dev)
    # ---- untimed: first calls (scratch allocation), then graph capture ----
    with torch.cuda.stream(stream):
        for b in sets:
            one_step(b)
    stream.synchronize()

    def capture(n_steps: int, first_set: int):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(n_steps):
                one_step(sets[(first_set + i) % S])
        return g

    use_graph = not args.no_graph
    G = ROLLOUT
    full_graph = capture(G, 0) if use_graph else None
    single = {}   # one-step graphs per state set, for the steps that do not fill a whole rollout

    def run_steps(n: int, start: int = 0) -> None:
        """Exactly n steps on `stream`; whole rollouts replay the 24-step graph, the remainder one-step graphs."""
        i = 0
        with torch.cuda.stream(stream):
            while i < n:
                k = (start + i) % S
                if use_graph and n - i >= G and k == 0:
                    full_graph.replay()
                    i += G
                elif use_graph:
                    if k not in single:
                        single[k] = capture(1, k)
                    single[k].replay()
                    i += 1
                else:
                    one_step(sets[k])
                    i += 1

    # warm-up (>= 3 steps); the one-step graphs the timed region may need are captured here, not inside it
    run_steps(W)
    if use_graph:
        for k in range(S):
            if k not in single:
                single[k] = capture(1, k)
    stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    # ---- timed region: exactly K steps ----
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    t_wall0 = time.perf_counter()
    # keep the device busy long enough for nvidia-smi to see it under load: the K steps are timed exactly once,
    ev0.record(stream)
    run_steps(K)
    ev1.record(stream)
    stream.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    value = world * N * K / (elapsed_ms * 1e-3)
    ms_per_step = elapsed_ms / K

    # ---- each step kernel alone for the roofline: same rotation, graph of G back-to-back launches ----
    def time_kernel(fn):
        with torch.cuda.stream(stream):
            for b in sets:
                fn(b)
        stream.synchronize()
        gk = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gk, stream=stream):
            for i in range(G):
                fn(sets[i % S])
        reps = max(3, min(200, K // G))
        with torch.cuda.stream(stream):
            for _ in range(3):
                gk.replay()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            for _ in range(reps):
                gk.replay()
            k1.record(stream)
        stream.synchronize()
        return 1e3 * k0.elapsed_time(k1) / (reps * G)

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    traffic_db = {}
    tp = ROOT / "profiles" / "traffic_latest.json"
    if tp.exists():
        try:
            traffic_db = json.loads(tp.read_text())
        except Exception:
            traffic_db = {}
    kernels = {}
    for kind, fn in (("pre_reset", lambda b: eng.step_pre_reset(b, **rng)), ("post_reset", lambda b: eng.step_post_reset(b, **rng))):
        us = time_kernel(fn)
        nbytes = spec.algorithmic_bytes_per_launch(kind) * N
        achieved = nbytes / (us * 1e-6) / 1e9
        kernels[kind] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic_db.get(kind, {}).get("dram_bytes_per_launch") if N == 4096 else None,
                         "kernel": f"mdp_step_kernel ({'DONES|REWARDS|COMPACT' if kind == 'pre_reset' else 'RESET|COMMAND|OBS'})",
                         "kernel_us": us, "bytes_per_launch": nbytes,
                         "bytes_per_env": spec.algorithmic_bytes_per_launch(kind), "peak_source": peak_src,
                         "kernel_share_of_step": us / (ms_per_step * 1e3)}
    dom = max(kernels, key=lambda k: kernels[k]["kernel_us"])
    roofline = dict(kernels[dom])
    roofline["other_kernel"] = kernels["post_reset" if dom == "pre_reset" else "pre_reset"]
    roofline["bytes_per_env_step_fused_accounting"] = spec.algorithmic_bytes_per_env_step()

    # ---- the neighbours of the path: each kernel alone, same rotation over the state sets ----
    neighbours = None
    if rank == 0 and world == 1 and not args.no_neighbours:
        neighbours = measure_neighbours(eng, spec, sets, N, time_kernel, peak, dev)

    # ---- the same step with an L2-resident working set (one state set, 14 MB): labelled separately ----
    l2_resident = None
    if use_graph and not args.no_e2e:
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            for _ in range(3):
                one_step(sets[0])
        stream.synchronize()
        with torch.cuda.graph(g1, stream=stream):
            for _ in range(G):
                one_step(sets[0])
        with torch.cuda.stream(stream):
            for _ in range(3):
                g1.replay()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps1 = max(3, min(100, K // G))
            r0.record(stream)
            for _ in range(reps1):
                g1.replay()
            r1.record(stream)
        stream.synchronize()
        ms1 = r0.elapsed_time(r1) / (reps1 * G)
        l2_resident = {"value": world * N * 1e3 / ms1, "unit": UNIT, "ms_per_step": ms1,
                       "note": "single state set (working set < L2); the headline value rotates over sets larger than L2"}

    # ---- e2e: same step through the C-ABI with HOST buffers (pinned), H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(eng, spec, sets, N, K, W, rank, world, dev, local_rank, one_step)

    # ---- multi-GPU hand-off (the one exchange of the path: rollout all-gather at the PPO boundary) ----
    handoff = None
    if world > 1 and not args.skip_handoff:
        handoff = measure_handoff(spec, N, world, dev, local_rank)

    # ---- CPU baseline (rank 0, N=1 run only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = time_cpu(spec, N, steps=10_000, warmup=2, budget_s=args.cpu_seconds)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{args.task} (BASELINE.json configs[2]), {N} envs/GPU, J={spec.J} B={spec.B} F={spec.Bt} "
                            f"R={spec.R} K={spec.K}, policy/critic rows {spec.obs[0].dim}/{spec.obs[1].dim}",
                "num_envs_per_gpu": N, "state_sets": S, "cuda_graph_steps": G if use_graph else 0,
                "pdl": args.pdl, "launch": eng.launch_config(),
                "l2_policy": f"rotating over {S} independent state sets (inputs+outputs+manager state "
                             f"{S * (sets[0].inputs.nbytes + sets[0].outputs.nbytes + sets[0].mdp.nbytes) / 1e6:.0f} MB > 126 MB L2)",
                "noise": "in-kernel Philox4x32-10 (0 bytes)", "parallelism": f"dp{world} (env shards, no data-path collective)",
            },
            "gpu_launches": launches_per_step * K,
            "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "handoff": handoff,
            "l2_resident": l2_resident, "neighbours": neighbours,
            "wall_s_timed_region": t_wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_neighbours(eng, spec, sets, N, time_kernel, peak, dev):
    """Side measurements (not part of `value`): the SURVEY.md 8(f) kernels that run around the step in a simulator
    loop - per physics sub-step the actuator model and the contact-sensor update, per env step the height scanner, per
    env step on pit terrains the command restriction. bytes = per-env tensors read + written (the height field and the
    terrain origins are shared, L2-resident tables)."""
    from robot_lab_b200 import terrain as terrain_host
    from robot_lab_b200.cfg import RayCasterCfg, TerrainCfg

    out = {}

    def add(name, fn, bytes_per_env, note):
        us = time_kernel(fn)
        gbs = bytes_per_env * N / (us * 1e-6) / 1e9
        out[name] = {"kernel_us": us, "bytes_per_env": bytes_per_env, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak,
                     "note": note}

    J, R, B, Bt, T = spec.J, spec.R, spec.B, spec.Bt, spec.T
    add("rl_actuator_step", lambda b: eng.actuator_step(b), 4 * 4 * J,
        "DCMotor: reads target, joint_pos, joint_vel, writes applied_torque; x decimation per env step")
    forces = torch.randn(N, B, 3, device=dev)
    add("rl_contact_sensor_update(ring)", lambda b: eng.contact_sensor_update(b, forces, 0.005, ring_slot=1),
        4 * (3 * B + 3 * B + 6 * Bt), "reads net_forces_w, writes one history slot, read-modify-writes 4 timers")
    if R > 0:
        ter = TerrainCfg()
        nx = int(round((ter.num_rows * ter.size[0] + 2 * ter.border_width) / ter.horizontal_scale)) + 1
        ny = int(round((ter.num_cols * ter.size[1] + 2 * ter.border_width) / ter.horizontal_scale)) + 1
        g = torch.Generator(device="cpu").manual_seed(7)
        heights = (torch.rand(nx // 8 + 2, ny // 8 + 2, generator=g) * 0.8)
        heights = torch.nn.functional.interpolate(heights[None, None], size=(nx, ny), mode="bilinear", align_corners=True)[0, 0]
        hf = terrain_host.HeightFieldBuffers(heights.contiguous().to(dev), -0.5 * (nx - 1) * ter.horizontal_scale,
                                             -0.5 * (ny - 1) * ter.horizontal_scale, ter.horizontal_scale,
                                             terrain_host.grid_pattern_ray_starts(RayCasterCfg()).to(dev))
        add("rl_height_scan_cast", lambda b: eng.height_scan_cast(b, hf), 4 * (7 + R + 1),
            f"{R} vertical rays per env over a {nx} x {ny} height field ({nx * ny * 4 / 1e6:.1f} MB, L2-resident gathers)")
    pit = TerrainCfg(sub_terrains=("pyramid_stairs", "pits", "boxes", "random_rough", "hf_pyramid_slope"),
                     proportions=(0.2, 0.15, 0.25, 0.3, 0.1))
    grid = terrain_host.TerrainGridBuffers.create(pit, "pits", dev)
    was = torch.zeros(N, dtype=torch.uint8, device=dev)
    add("rl_command_pit_restrict", lambda b: eng.command_pit_restrict(b, grid, was, seed=1, use_random_inputs=False),
        4 * 2 + 2, "argmin over 200 terrain origins per env (shared memory) + command rewrite for pit envs")
    return out


def measure_e2e(eng, spec, sets, N, K, W, rank, world, dev, local_rank, one_step):
    """Every step: one H2D copy of that step's inputs (physics/sensor state + policy action) from pinned host
    memory, the three launches, one D2H read of the step's results (observation rows, reward, done masks).
    Three streams pipeline copy-in / compute / copy-out over a ring of device sets."""
    import torch.distributed as dist

    ring = min(4, len(sets))
    dsets = sets[:ring]
    in_bytes, out_bytes = dsets[0].inputs.nbytes, dsets[0].outputs.nbytes
    # first-touch the pinned staging buffers on the NUMA node the GPU hangs off (a remote node costs up to 40 % of
    # the PCIe rate on a two-socket host); the affinity is restored right after the allocation
    numa = None
    old_aff = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    try:
        props = torch.cuda.get_device_properties(dev)
        if all(hasattr(props, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = f"{int(props.pci_domain_id):04x}:{int(props.pci_bus_id):02x}:{int(props.pci_device_id):02x}.0"
        else:
            import pynvml
            pynvml.nvmlInit()
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
            bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
            if len(bdf.split(":")[0]) == 8:
                bdf = bdf[4:]
        cpus = Path(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read_text().strip()
        want = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            want.update(range(int(lo), int(hi or lo) + 1))
        want &= old_aff
        if want:
            os.sched_setaffinity(0, want)
            numa = f"pinned buffers first-touched on the GPU-local CPUs ({cpus})"
    except Exception as e:   # not fatal: the buffers land wherever the allocator puts them
        numa = f"no NUMA placement ({type(e).__name__})"
    host_in = [torch.empty(in_bytes, dtype=torch.uint8).pin_memory() for _ in range(ring)]
    host_out = [torch.empty(out_bytes, dtype=torch.uint8).pin_memory() for _ in range(ring)]
    for h, b in zip(host_in, dsets):
        h.copy_(b.inputs.buf[:in_bytes].cpu())
    for h in host_out:
        h.zero_()
    if old_aff is not None:
        try:
            os.sched_setaffinity(0, old_aff)
        except Exception:
            pass
    s_in, s_cmp, s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
    ev_in = [torch.cuda.Event() for _ in range(ring)]
    ev_cmp = [torch.cuda.Event() for _ in range(ring)]
    ev_out = [torch.cuda.Event() for _ in range(ring)]
    steps = max(200, min(K, 600))   # its own step count: long enough for the 4-deep pipeline to reach steady state

    # the three launches of a set, captured once through the C-ABI calls: replaying them keeps the host side of a
    # step at a handful of stream operations (the Python / ctypes cost of three rl_* calls would otherwise bound it)
    graphs = []
    with torch.cuda.stream(s_cmp):
        for b in dsets:
            one_step(b)
    torch.cuda.synchronize(dev)
    for b in dsets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s_cmp):
            one_step(b)
        graphs.append(g)

    def run(n):
        for i in range(n):
            r = i % ring
            b = dsets[r]
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_cmp[r])          # the set's previous compute has consumed its inputs
                b.inputs.buf[:in_bytes].copy_(host_in[r], non_blocking=True)
                ev_in[r].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(ev_in[r])
                s_cmp.wait_event(ev_out[r])         # the set's previous results have been read back
                graphs[r].replay()
                ev_cmp[r].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[r])
                host_out[r].copy_(b.outputs.buf[:out_bytes], non_blocking=True)
                ev_out[r].record(s_out)

    run(max(24, min(W, 48)))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier(device_ids=[local_rank])
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(s_in)
    run(steps)
    for s in (s_in, s_cmp):
        s_out.wait_stream(s)
    t1.record(s_out)
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": world * N * steps / (ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": dsets[0].input_bytes(), "d2h_bytes_per_step": dsets[0].output_bytes(),
            "steps": steps, "ms_per_step": ms / steps, "numa": numa,
            "path": "pinned host buffers -> 1 H2D -> rl_process_action + 2 x rl_step (C-ABI calls captured once, replayed as a CUDA graph) -> 1 D2H per step, 3-stream pipeline over a ring of 4 device sets"}


def measure_handoff(spec, N, world, dev, local_rank):
    """NCCL all-gather of one rank-local rollout buffer (24 steps x N envs x 320 fp32) - the single exchange the
    north star places at the PPO hand-off. Reported beside the step throughput, not inside it."""
    import torch.distributed as dist

    from robot_lab_b200.rollout import rollout_row_width

    width = rollout_row_width(spec)
    local = torch.randn(ROLLOUT, N, width, device=dev)
    gathered = torch.empty(world * ROLLOUT, N, width, device=dev)
    for _ in range(3):
        dist.all_gather_into_tensor(gathered, local)
    torch.cuda.synchronize(dev)
    dist.barrier(device_ids=[local_rank])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        dist.all_gather_into_tensor(gathered, local)
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    nbytes = local.numel() * 4
    return {"collective": "ncclAllGather (torch.distributed, NVLink/NVSwitch)", "bytes_per_rank": nbytes,
            "ms": ms, "bus_gbs": nbytes * (world - 1) / (ms * 1e-3) / 1e9, "per_rollout_steps": ROLLOUT,
            "amortised_ms_per_step": ms / ROLLOUT}


if __name__ == "__main__":
    main()

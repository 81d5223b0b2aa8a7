"""Where a launch of the new step kernels goes (build variant "stamps": python -m robot_lab_b200.build --variant stamps;
run with RL_MDPSTEP_LIB=robot_lab_b200/_lib/libmdpstep_stamps.so). Per CTA clock64 stamps: 1 entry, 2 loads issued +
prologue done, 3 record resident, 4 tasks start, 16+w end of warp w's tasks, 5 after the task barrier, 6 tail done;
globaltimer (ns) at entry / end of every CTA gives the launch's spread over the grid.

Usage (GPU box): python tools/v2_timeline.py [num_envs] [task_key]
"""
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))

import torch  # noqa: E402

import helpers as H  # noqa: E402
from robot_lab_b200.engine import MdpStepEngine  # noqa: E402
from robot_lab_b200.synthetic import make_state  # noqa: E402

WARM = "--warm" in sys.argv   # stamp the 4th of four identical back-to-back launches (same state set, nothing in between)
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(argv[0]) if len(argv) > 0 else 4096
key = argv[1] if len(argv) > 1 else "go2_rough"
cfg, spec = H.make_spec(key)
eng = MdpStepEngine(spec, "cuda:0")
cc = eng.cluster_config(N)
nw = cc["warps_per_cta"]
sets = []
for i in range(6):
    b = eng.new_buffers(N)
    b.load_logical(make_state(spec, N, seed=1234 + i))
    b.cmd_uniforms, b.obs_uniforms = None, [None, None]
    sets.append(b)
rng = dict(use_random_inputs=False, use_step_counter=True)
grid = (N // (32 * cc["tiles_per_cta"])) * cc["cluster_size"]
dbg = torch.zeros(grid, 64, dtype=torch.int64, device="cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")
print(f"{key} N={N} config {cc} grid {grid} {'WARM (4th identical launch)' if WARM else 'cold (L2 flushed, other kernels in between)'}")
for kind in ("pre", "post"):
    rows = []
    for rep in range(8):
        b = sets[rep % len(sets)]
        eng.set_debug_buffer(None)
        eng.process_action(b)
        if kind == "post":
            eng.step_pre_reset(b, **rng)
        fn = eng.step_pre_reset if kind == "pre" else eng.step_post_reset
        if WARM:
            for _ in range(3):
                fn(b, **rng)
        else:
            flush.fill_(rep)                  # the launch reads HBM
        torch.cuda.synchronize()
        dbg.zero_()
        eng.set_debug_buffer(dbg)
        fn(b, **rng)
        torch.cuda.synchronize()
        eng.set_debug_buffer(None)
        if kind == "pre":
            eng.step_post_reset(b, **rng)
        d = dbg.cpu()
        if rep >= 2:
            rows.append(d)
    med = lambda xs: statistics.median(xs)
    out = {}
    if kind == "pre":
        for name, a_, b_ in (("  entry -> mbarrier armed", 1, 9), ("  first CTA barrier", 9, 10), ("  this warp's copies issued", 10, 11),
                             ("  norm prepass + joint constants", 11, 12), ("  zero slots + second barrier", 12, 2)):
            out[name] = med([float((d[:, b_] - d[:, a_]).float().median()) for d in rows])
    for name, a_, b_ in (("prologue + load issue", 1, 2), ("wait for the record", 2, 3), ("reset / prepass", 3, 4),
                         ("tasks + barrier", 4, 5), ("tail", 5, 6), ("entry -> tail done", 1, 6)):
        out[name] = med([float((d[:, b_] - d[:, a_]).float().median()) for d in rows])
    task_end = [d[:, 16:16 + nw] - d[:, 4:5] for d in rows]
    out["slowest warp's tasks"] = med([float(t.max(dim=1).values.float().median()) for t in task_end])
    out["fastest warp's tasks"] = med([float(t.min(dim=1).values.float().median()) for t in task_end])
    per_warp = torch.stack([t.float().median(dim=0).values for t in task_end]).median(dim=0).values
    span = med([float((d[:, 7].max() - d[:, 0].min())) for d in rows])
    first_last_entry = med([float((d[:, 0].max() - d[:, 0].min())) for d in rows])
    print(f"--- {kind}: cycles per CTA (median over CTAs, then over launches); SM clock ~1.9 GHz")
    for k, v in out.items():
        print(f"   {k:28s} {v:9.0f} cycles  {v / 1.9e3:6.2f} us")
    print("   per-warp task cycles:", " ".join(f"{int(x)}" for x in per_warp.tolist()))
    print(f"   globaltimer: first CTA entry -> last CTA tail {span / 1e3:.2f} us; entry spread over the grid {first_last_entry / 1e3:.2f} us")
    # per-task cycles (stamps build: slot 32 + task index, in schedule order)
    if kind == "pre":
        names = ["DONES"] + [r.name for r in spec.rewards if r.weight != 0.0 and r.type_name != "is_terminated"]
    else:
        names = ["COMMAND"] + [f"LOG part {p}" for p in range(4)]
        for g, grp in enumerate(spec.obs):
            for t in grp.terms:
                if t.name in ("velocity_commands", "height_scan") or t.type_name in ("generated_commands", "height_scan"):
                    continue
                for lo in range(0, t.dim, 32):
                    names.append(f"OBS {grp.name}.{t.name}[{lo}:{min(lo + 32, t.dim)}]")
    tcy = torch.stack([d[:, 32:64].float().median(dim=0).values for d in rows]).median(dim=0).values.tolist()
    print("   per-task cycles (median over CTAs): " + ", ".join(f"{names[i] if i < len(names) else i}={int(c)}" for i, c in enumerate(tcy) if c > 0))
    if kind == "post":   # the warp that finished the logging reduction of the launch (one-level form: stamps 56 - 58)
        fin = [int(torch.argmax(d[:, 58])) for d in rows]
        if all(int(d[c_, 58]) > 0 for d, c_ in zip(rows, fin)):
            tl = lambda a_, b_: med([float(d[c_, b_] - d[c_, a_]) for d, c_ in zip(rows, fin)]) / 1e3
            tl0 = lambda b_: med([float(d[c_, b_] - d[:, 0].min()) for d, c_ in zip(rows, fin)]) / 1e3
            print(f"   logging finisher CTA (one-level form; globaltimer, us after the first entry): starts {tl0(56):.2f}, rows staged +{tl(56, 57):.2f}, "
                  f"means written +{tl(57, 58):.2f}; its CTA ends {tl0(7):.2f}")
    ends = [torch.sort(d[:, 7] - d[:, 0].min()).values.float() for d in rows]
    q = lambda f: med([float(e_[int(f * (len(e_) - 1))]) for e_ in ends]) / 1e3
    print(f"   CTA end times (us after the first entry; the launches have no launch-wide tail): 50 % {q(0.5):.2f}, 90 % {q(0.9):.2f}, "
          f"99 % {q(0.99):.2f}, last {q(1.0):.2f}")


# ---- launch-overhead probes (stamps build): the same launches, returning at entry / right after the record is resident ----
import ctypes as C  # noqa: E402


def graph_time(fn, n=48, reps=20):
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        for i in range(4):
            fn(sets[i % len(sets)])
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for i in range(n):
                fn(sets[i % len(sets)])
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            g.replay()
        e1.record(st)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n)


def set_probe(code):
    eng.lib.rl_ctx_set_debug_buffer(eng._ctx, C.c_void_p(code) if code else None)


print("--- graph of 48 back-to-back launches, us per launch: whole kernel | returns after the record is resident | returns at entry")
for kind, fn in (("pre", lambda b: eng.step_pre_reset(b, **rng)), ("post", lambda b: eng.step_post_reset(b, **rng)),
                 ("process_action", lambda b: eng.process_action(b))):
    row = []
    for code in (0, 2, 1):
        set_probe(code)
        row.append(graph_time(fn))
    set_probe(0)
    print(f"   {kind:16s} {row[0]:7.2f} | {row[1]:7.2f} | {row[2]:7.2f}")

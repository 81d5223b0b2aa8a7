"""Diagnostic: run the fused step on the GPU for every in-scope task and print the error against the oracle.

Usage (on a GPU box): python tools/parity_report.py [num_envs]
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))

import torch  # noqa: E402

import helpers as H  # noqa: E402
from oracle import mdp_port as port  # noqa: E402
from robot_lab_b200 import _native as nat  # noqa: E402
from robot_lab_b200.engine import MdpStepEngine  # noqa: E402
from robot_lab_b200.synthetic import make_state  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for key in H.TASKS:
    cfg, spec = H.make_spec(key)
    st = make_state(spec, N)
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(N)
    b.load_logical(st)
    eng.step(b)
    torch.cuda.synchronize()
    got, ref = H.gpu_step_outputs(b), H.oracle_step(spec, st)
    print(f"== {key} N={N} resets={len(ref['reset_ids'])}")
    for k in ref:
        g, r = got[k].float(), ref[k].float()
        if g.shape != r.shape:
            print(f"   {k:22s} SHAPE {tuple(g.shape)} vs {tuple(r.shape)}")
            continue
        both_inf = torch.isinf(g) & torch.isinf(r)
        err = torch.where(both_inf, torch.zeros_like(g), (g - r).abs())
        rel = err / r.abs().clamp(min=1e-6)
        nbad = int((err > 1e-6 + 1e-5 * r.abs()).sum())
        exact = bool(torch.equal(g, r))
        print(f"   {k:22s} max_abs {float(err.max()) if err.numel() else 0:.3e} max_rel {float(rel.max()) if rel.numel() else 0:.3e} bad {nbad} bit_exact {exact}")
    # per-term (reload the inputs: the step above advanced the manager state)
    b.load_logical(st)
    d = port.Derived({**st, "terminated": ref["terminated"]}, spec)
    for t in spec.rewards:
        v = eng.term_eval(t, b, terminated=b.terminated)
        torch.cuda.synchronize()
        r = port.reward_term(t, {**st, "terminated": ref["terminated"]}, spec, d)
        err = (v.cpu() - r).abs()
        print(f"      term {t.name:28s} max_abs {float(err.max()):.3e} max_ref {float(r.abs().max()):.3e} bit_exact {bool(torch.equal(v.cpu(), r))}")
    eng.close()

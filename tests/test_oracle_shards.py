"""envs shard by contiguous id ranges with NO data-path collective. At the oracle level: the
full MDP step of rank r over envs [r*N, (r+1)*N) with env_id_offset = r*N (production Philox streams: command resample,
observation noise) equals the corresponding slice of one process stepping all envs - every output, bit for bit. The
CUDA path is tied to the same counters by tests/test_gpu_step_parity.py::test_production_philox_stream_matches_oracle."""

import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200.synthetic import make_state


def _slice(st, lo, hi):
    out = {}
    for k, v in st.items():
        if k in ("cmd_uniforms",):
            out[k] = v[:, lo:hi].clone()
        else:
            out[k] = v[lo:hi].clone()
    return out


def test_two_shards_reproduce_the_single_process_step_bit_for_bit():
    cfg, spec = H.make_spec("go2_rough")
    n, half = 256, 128
    st = make_state(spec, n, seed=77)
    seed, step = 0xC0FFEE, 17
    full = port.step(spec, st, {"seed": seed, "step": step, "env_id_offset": 0})
    parts = [port.step(spec, _slice(st, r * half, (r + 1) * half), {"seed": seed, "step": step, "env_id_offset": r * half})
             for r in range(2)]
    assert len(full["reset_ids"]) > 0
    for k, v in full.items():
        if k == "reset_ids":
            cat = torch.cat([p[k] + r * half for r, p in enumerate(parts)])     # local ids + the shard's offset
        else:
            cat = torch.cat([p[k] for p in parts])
        assert cat.shape == v.shape, k
        if v.dtype.is_floating_point:
            assert torch.equal(cat.view(torch.int32), v.view(torch.int32)), k
        else:
            assert torch.equal(cat, v), k

"""Multi-GPU hand-off logic on CPU: 2 ranks over gloo all-gather their rollout shards; the result equals the
single-process concatenation in global env-id order."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, steps, n_local, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robot_lab_b200.rollout import RolloutBuffer, shard_range

    cfg, spec = H.make_spec("go2_rough")
    lo, hi = shard_range(world * n_local, rank, world)
    assert (lo, hi) == (rank * n_local, (rank + 1) * n_local)
    buf = RolloutBuffer(spec, n_local, steps, "cpu")
    g = torch.Generator().manual_seed(0)
    full = torch.randn(steps, world * n_local, buf.width, generator=g)
    for t in range(steps):
        row = full[t, lo:hi]
        s = buf.slices
        buf.add(row[:, s["obs_policy"]], row[:, s["obs_critic"]], row[:, s["action"]], row[:, s["reward"]].squeeze(-1),
                row[:, s["done"]].squeeze(-1), row[:, s["value"]], row[:, s["log_prob"]], row[:, s["mean"]], row[:, s["sigma"]])
    gathered = buf.all_gather()
    ok = torch.equal(gathered, full)
    torch.save(torch.tensor([int(ok), gathered.shape[1]]), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_rollout_all_gather(tmp_path):
    world, steps, n_local = 2, 3, 16
    mp.spawn(_worker, args=(world, _free_port(), steps, n_local, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok, n = torch.load(tmp_path / f"r{r}.pt").tolist()
        assert ok == 1 and n == world * n_local


def test_shard_ranges_cover_everything():
    from robot_lab_b200.rollout import rollout_row_width, shard_range

    for total, world in ((32768, 8), (4097, 8), (7, 3)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    cfg, spec = H.make_spec("go2_rough")
    assert rollout_row_width(spec) == 320  # 45+235+12+1+1+1+1+12+12

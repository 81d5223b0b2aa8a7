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


def _storage_worker(rank, world, port, steps, n_local, out_dir, every=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robot_lab_b200.rollout import RolloutStorage

    cfg, spec = H.make_spec("go2_rough")
    store = RolloutStorage(spec, n_local, steps, "cpu", gather_every=every)
    g = torch.Generator().manual_seed(1)
    names = ("obs_policy", "obs_critic", "action", "mean", "sigma", "reward", "value", "log_prob")
    full = {n: torch.randn(steps, world * n_local, *( (store._planes[n][1],) if store._planes[n][1] > 1 else ()), generator=g) for n in names}
    done = torch.rand(steps, world * n_local, generator=g) < 0.3
    lo, hi = rank * n_local, (rank + 1) * n_local
    ok = True
    for t in range(steps):     # "the step writes slab t", then its gather is issued at once
        for n in names:
            store.plane(n, t).copy_(full[n][t, lo:hi])
        term, trunc = store.done_bytes(t)
        term.copy_(done[t, lo:hi].to(torch.uint8))
        trunc.zero_()
        store.gather_step(t)
    store.finish()
    for t in range(steps):
        for n in names:
            ok = ok and torch.equal(store.global_plane(n, t), full[n][t])
        got = torch.cat([store.plane("done", t, store.gathered[:, r].contiguous()).view(torch.uint8)[:n_local] for r in range(world)])
        ok = ok and torch.equal(got.bool(), done[t])
    whole = store.gather_all()
    ok = ok and torch.equal(whole[rank], store.data.view(-1)) and whole.shape == (world, steps * n_local * store.width)
    torch.save(torch.tensor([int(ok)]), os.path.join(out_dir, f"s{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("every", [1, 2])
def test_two_rank_rollout_storage_streamed_gather(tmp_path, every):
    """RolloutStorage: the step's results live in per-step slabs (written in place), every chunk of slabs is gathered as
    soon as it is complete, and the gathered planes are the single-process tensors in global env-id order."""
    world, steps, n_local = 2, 4, 32
    mp.spawn(_storage_worker, args=(world, _free_port(), steps, n_local, str(tmp_path), every), nprocs=world, join=True)
    for r in range(world):
        assert torch.load(tmp_path / f"s{r}.pt").item() == 1


def test_rollout_storage_layout_single_process():
    from robot_lab_b200.rollout import RolloutStorage, rollout_row_width

    cfg, spec = H.make_spec("go2_rough")
    store = RolloutStorage(spec, 64, 3, "cpu")
    assert store.data.shape == (3, 64 * rollout_row_width(spec))
    # planes tile the slab without gaps or overlap, every plane starts 128-byte aligned (N is a multiple of 32)
    spans = sorted((o, o + 64 * w) for o, w in store._planes.values())
    assert spans[0][0] == 0 and spans[-1][1] == 64 * store.width and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert all((o * 4) % 128 == 0 for o, _ in spans)
    assert store.plane("obs_critic", 1).shape == (64, spec.obs[1].dim) and store.plane("reward", 2).shape == (64,)
    term, trunc = store.done_bytes(0)
    assert term.dtype == torch.uint8 and term.shape == (64,) and trunc.data_ptr() == term.data_ptr() + 64


def test_shard_ranges_cover_everything():
    from robot_lab_b200.rollout import rollout_row_width, shard_range

    for total, world in ((32768, 8), (4097, 8), (7, 3)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    cfg, spec = H.make_spec("go2_rough")
    assert rollout_row_width(spec) == 320  # 45+235+12+1+1+1+1+12+12

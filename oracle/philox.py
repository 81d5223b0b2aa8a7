"""Philox4x32-10 in numpy - the oracle's copy of the kernel's counter-based random streams.

TEST INFRASTRUCTURE (see oracle/mdp_port.py). The algorithm is the published one (Salmon, Moraes, Dror, Shaw:
"Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 ``philox4x32_R(10, ...)``); it is pinned by the
Random123 known-answer vectors in ``tests/test_philox.py``. The counter layout mirrors
``rl_philox`` in robot_lab_b200/csrc/mdp_step.cu:

    counter = (env_lo, step_lo, step_hi ^ env_hi, stream << 16 | block)     key = (seed_lo, seed_hi)
    uniform = (x >> 8) * 2**-24   in [0, 1)

The reference draws its noise with torch.rand_like / Tensor.uniform_ (V/velocity_env_cfg.py:140-187 noise cfgs,
UniformVelocityCommand._resample_command [IL]); torch's generator offsets cannot be reproduced inside a fused
kernel, so bit-level parity tests feed the uniforms as inputs and this module checks the production stream.
"""

from __future__ import annotations

import numpy as np
import torch

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_COMMAND = 1
STREAM_RESET_COMMAND = 2
STREAM_OBS = 16
MAX_OBS_TERMS = 12


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [..., 4] uint32, key [..., 2] uint32 -> [..., 4] uint32."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32).copy()
    k1 = key[..., 1].astype(np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0]
            p1 = M1 * c[2]
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
            c = [hi1 ^ c[1] ^ k0.astype(np.uint64), lo1, hi0 ^ c[3] ^ k1.astype(np.uint64), lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def u01(x: np.ndarray) -> np.ndarray:
    return ((x >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def draw(env_ids: np.ndarray, seed: int, step: int, env_id_offset: int, stream: int, block: int | np.ndarray) -> np.ndarray:
    """[len(env_ids), 4] uniforms of one (stream, block) for the given local env ids."""
    genv = env_ids.astype(np.int64) + np.int64(env_id_offset)
    genv = genv.astype(np.uint64)
    n = genv.shape[0]
    ctr = np.zeros((n, 4), dtype=np.uint32)
    ctr[:, 0] = (genv & MASK32).astype(np.uint32)
    ctr[:, 1] = np.uint32(step & 0xFFFFFFFF)
    ctr[:, 2] = np.uint32((step >> 32) & 0xFFFFFFFF) ^ (genv >> np.uint64(32)).astype(np.uint32)
    ctr[:, 3] = (np.uint32(stream) << np.uint32(16)) | np.asarray(block, dtype=np.uint32)
    key = np.zeros((n, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    return u01(philox4x32_10(ctr, key))


def command_uniforms(n: int, seed: int, step: int, env_id_offset: int, stream: int) -> torch.Tensor:
    ids = np.arange(n)
    a = draw(ids, seed, step, env_id_offset, stream, 0)
    b = draw(ids, seed, step, env_id_offset, stream, 1)
    u = np.concatenate([a, b[:, :3]], axis=1).T.copy()  # [7, N]
    return torch.from_numpy(u)


def obs_uniforms(n: int, dims: list[int], group: int, seed: int, step: int, env_id_offset: int) -> torch.Tensor:
    ids = np.arange(n)
    cols = []
    for ti, dim in enumerate(dims):
        stream = STREAM_OBS + group * MAX_OBS_TERMS + ti
        quads = [draw(ids, seed, step, env_id_offset, stream, q) for q in range((dim + 3) // 4)]
        if quads:
            cols.append(np.concatenate(quads, axis=1)[:, :dim])
    out = np.concatenate(cols, axis=1) if cols else np.zeros((n, 0), dtype=np.float32)
    return torch.from_numpy(np.ascontiguousarray(out))

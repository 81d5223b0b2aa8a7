"""Philox4x32-10: the oracle's numpy implementation against the Random123 known-answer vectors
(kat_vectors: philox4x32 10 rounds), and the uniform conversion / counter layout used by the kernel."""

import numpy as np

from oracle import philox

KAT = [  # (counter, key, expected) - Random123 kat_vectors "philox4x32 10"
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF),
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_known_answer_vectors():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert tuple(int(x) for x in got) == want


def test_uniforms_are_in_unit_interval_and_streams_differ():
    u = philox.command_uniforms(4096, seed=7, step=3, env_id_offset=0, stream=philox.STREAM_COMMAND).numpy()
    assert u.shape == (7, 4096) and u.dtype == np.float32
    assert (u >= 0).all() and (u < 1).all()
    assert abs(u.mean() - 0.5) < 0.01
    v = philox.command_uniforms(4096, seed=7, step=3, env_id_offset=0, stream=philox.STREAM_RESET_COMMAND).numpy()
    assert not np.array_equal(u, v)
    w = philox.command_uniforms(4096, seed=7, step=4, env_id_offset=0, stream=philox.STREAM_COMMAND).numpy()
    assert not np.array_equal(u, w)


def test_env_id_offset_shards_consistently():
    """Rank r with env_id_offset r*N draws what a single process draws for global envs [r*N, (r+1)*N)."""
    full = philox.obs_uniforms(64, [3, 12, 187], 1, seed=9, step=5, env_id_offset=0).numpy()
    lo = philox.obs_uniforms(32, [3, 12, 187], 1, seed=9, step=5, env_id_offset=0).numpy()
    hi = philox.obs_uniforms(32, [3, 12, 187], 1, seed=9, step=5, env_id_offset=32).numpy()
    assert np.array_equal(full[:32], lo) and np.array_equal(full[32:], hi)

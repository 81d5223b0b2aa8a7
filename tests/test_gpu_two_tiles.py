"""Two tiles per CTA (rl_ctx_set_launch_config(ctx, 64, 16)): warps w and w + 16 of a 1024-thread CTA run the same
task on neighbouring 32-env tiles, each tile with its own record, mbarrier and named barrier. A tile executes exactly
the code a one-tile CTA executes, so every output must be BIT-identical to the one-tile launch - for full steps, the
two-launch env step, ragged / odd tile counts, and env-id lists - and match the oracle to the usual bar.

The configuration is a measured dead end at the sizes of BASELINE.json (profiles/r1_summary.md section 8: 12 - 26 %
slower per launch than one tile per CTA) and stays an opt-in knob; the tests keep its kernel instantiations honest.
"""

import pytest
import torch

import helpers as H
from robot_lab_b200 import _native as nat
from robot_lab_b200.synthetic import make_state

pytestmark = pytest.mark.gpu


def _engine(spec, envs_per_cta):
    from robot_lab_b200.engine import MdpStepEngine

    eng = MdpStepEngine(spec, "cuda:0")
    eng.set_launch_config(16, envs_per_cta)
    return eng


def _outputs(b):
    out = H.gpu_step_outputs(b)
    for k in ("heading_target", "time_left", "is_heading_env", "is_standing_env", "metric_error_vel_xy",
              "metric_error_vel_yaw", "action", "prev_action"):
        out[k] = b.logical(k).cpu().contiguous()
    out["log_episode_sum_mean"] = b.log_episode_sum_mean.cpu().clone()
    out["log_done_term_count"] = b.log_done_term_count.cpu().clone()
    out["log_metric_mean"] = b.log_metric_mean.cpu().clone()
    return out


def _assert_bit_equal(a, b):
    assert a.keys() == b.keys()
    for k in a:
        x, y = a[k], b[k]
        assert x.shape == y.shape, k
        if x.dtype.is_floating_point:   # bitwise, NaN-safe
            assert torch.equal(x.contiguous().view(torch.int32), y.contiguous().view(torch.int32)), k
        else:
            assert torch.equal(x, y), k


@pytest.mark.parametrize("key,n", [("go2_rough", 4096), ("go2_rough", 1), ("go2_rough", 33), ("go2_rough", 65),
                                   ("go2_rough", 4097), ("go2_flat", 4096), ("a1_flat", 64), ("a1_flat", 95)])
def test_fused_step_two_tiles_equals_one_tile_and_oracle(native_lib, key, n):
    cfg, spec = H.make_spec(key)
    st = make_state(spec, n)
    outs = []
    for epc in (32, 64):
        eng = _engine(spec, epc)
        b = eng.new_buffers(n)
        b.load_logical(st)
        eng.step(b)
        torch.cuda.synchronize()
        outs.append(_outputs(b))
        eng.close()
    _assert_bit_equal(outs[0], outs[1])
    H.compare_outputs(outs[1], H.oracle_step(spec, st))   # iterates the oracle's keys


@pytest.mark.parametrize("n", [2048, 2049 + 32, 100])
def test_two_launch_env_step_two_tiles_equals_one_tile(native_lib, n):
    """process_action -> DONES|REWARDS|COMPACT -> RESET|COMMAND|OBS, three env steps in a row, in-kernel Philox."""
    cfg, spec = H.make_spec("go2_rough")
    st = make_state(spec, n)
    outs = []
    for epc in (32, 64):
        eng = _engine(spec, epc)
        b = eng.new_buffers(n)
        b.load_logical(st)
        b.cmd_uniforms, b.obs_uniforms = None, [None, None]
        for _ in range(3):
            eng.process_action(b)
            eng.step_pre_reset(b, use_random_inputs=False, use_step_counter=True)
            eng.step_post_reset(b, use_random_inputs=False, use_step_counter=True)
        torch.cuda.synchronize()
        outs.append(_outputs(b))
        eng.close()
    assert int(outs[0]["reset_ids"].numel()) > 0 or n < 200
    _assert_bit_equal(outs[0], outs[1])


def test_env_id_list_launches_two_tiles_equals_one_tile(native_lib):
    """Fused step with SKIP_DONE_ENVS, manager reset and COMMAND|OBS refresh on the reset-id list (gathered tiles,
    device-side count): the id-list launches run through the same two-tile kernel."""
    cfg, spec = H.make_spec("go2_rough")
    n = 3000
    st = make_state(spec, n)
    outs = []
    for epc in (32, 64):
        eng = _engine(spec, epc)
        b = eng.new_buffers(n)
        b.load_logical(st)
        eng.step(b, phases=nat.PHASE_ALL | nat.PHASE_SKIP_DONE_ENVS)
        eng.reset_envs(b, b.reset_ids, b.n_reset)
        eng.step(b, phases=nat.PHASE_COMMAND | nat.PHASE_OBS, env_ids=b.reset_ids, n_env_ids=b.n_reset)
        torch.cuda.synchronize()
        outs.append(_outputs(b))
        eng.close()
    _assert_bit_equal(outs[0], outs[1])


def test_two_tiles_refused_where_the_records_do_not_fit(native_lib):
    """G1 rough: one record is larger than half an SM's shared memory (default build); 8 warps per tile and the generic
    kernel have no two-tile form."""
    import ctypes as C

    cfg, spec = H.make_spec("g1_rough")
    from robot_lab_b200.engine import MdpStepEngine

    eng = MdpStepEngine(spec, "cuda:0")
    cs = spec.to_ctypes()
    record = native_lib.rl_tile_record_bytes(C.byref(cs))
    if 2 * ((record + 127) // 128 * 128) > 232448:          # default build: 117 120 B per record
        with pytest.raises(nat.NativeError):
            eng.set_launch_config(16, 64)
    else:                                                  # a build variant with a smaller record may fit two
        eng.set_launch_config(16, 64)
    with pytest.raises(nat.NativeError):
        eng.set_launch_config(8, 64)
    eng.set_launch_config(16, 32)
    eng.close()
    cfg, spec = H.make_spec("go2_rough", full_layout=True)   # not a baked spec -> generic kernel
    eng = MdpStepEngine(spec, "cuda:0")
    with pytest.raises(nat.NativeError):
        eng.set_launch_config(16, 64)
    eng.close()

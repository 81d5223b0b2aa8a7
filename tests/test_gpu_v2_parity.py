"""The cluster kernels of the two env-step launches (csrc/mdp_step_v2.cu) against the general kernel (csrc/mdp_step.cu):
same term functions, same operand order -> every output BIT-IDENTICAL, for every compiled (cluster size x tiles per CTA)
configuration, every baked task, noise-as-input and production Philox streams, several chained env steps. The general
kernel itself is checked against the CPU oracle in test_gpu_step_parity.py; test_v2_two_launch_step_matches_oracle closes
the triangle directly.
"""

import os

import pytest
import torch

import helpers as H
from robot_lab_b200 import _native as nat
from robot_lab_b200.synthetic import make_state

pytestmark = pytest.mark.gpu

CONFIGS = ["1x1x16", "1x1x8", "1x1x4", "2x2x16", "4x4x16"]   # cluster size x tiles per CTA x warps per CTA
FIELDS = ("command", "heading_target", "time_left", "is_heading_env", "is_standing_env", "metric_error_vel_xy",
          "metric_error_vel_yaw", "episode_length", "episode_sums", "action", "prev_action", "step_reward", "joint_target")


def _engine(spec, v2_cfg):
    """v2_cfg: None = general kernel only, else 'CxG'."""
    from robot_lab_b200.engine import MdpStepEngine

    old = {k: os.environ.get(k) for k in ("RL_MDPSTEP_V2", "RL_MDPSTEP_V2_CFG")}
    try:
        if v2_cfg is None:
            os.environ["RL_MDPSTEP_V2"] = "0"
            os.environ.pop("RL_MDPSTEP_V2_CFG", None)
        else:
            os.environ.pop("RL_MDPSTEP_V2", None)
            os.environ["RL_MDPSTEP_V2_CFG"] = v2_cfg
        return MdpStepEngine(spec, "cuda:0")   # the switches are read at rl_ctx_create
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _snapshot(b) -> dict:
    out = {k: b.t[k].clone() for k in FIELDS}
    n = int(b.n_reset.item())
    out.update(reward=b.reward.clone(), terminated=b.terminated.clone(), truncated=b.truncated.clone(),
               done_bits=b.done_bits.clone(), n_reset=b.n_reset.clone(), reset_ids=b.reset_ids[:n].clone(),
               log_sum=b.log_episode_sum_mean.clone(), log_done=b.log_done_term_count.clone(), log_metric=b.log_metric_mean.clone())
    for g in (0, 1):
        if b.obs[g] is not None:
            out[f"obs{g}"] = b.obs[g].clone()
    return out


def _assert_identical(a: dict, b: dict, where: str):
    bad = []
    for k in a:
        x, y = a[k], b[k]
        same = x.shape == y.shape and torch.equal(x.view(torch.uint8) if x.dtype != torch.bool else x, y.view(torch.uint8) if y.dtype != torch.bool else y)
        if not same:
            n = int((x != y).sum()) if x.shape == y.shape else -1
            bad.append(f"{k} ({n} elements differ)")
    assert not bad, f"{where}: cluster kernel != general kernel in: " + ", ".join(bad)


def _three_steps(key, n, v2_cfg, philox, steps=3):
    cfg, spec = H.make_spec(key)
    eng_a, eng_b = _engine(spec, None), _engine(spec, v2_cfg)
    assert eng_a.cluster_config(n)["cluster_size"] == 0
    cc = eng_b.cluster_config(n)
    assert f"{cc['cluster_size']}x{cc['tiles_per_cta']}x{cc['warps_per_cta']}" == v2_cfg
    ba, bb = eng_a.new_buffers(n), eng_b.new_buffers(n)
    rng = dict(seed=7, env_id_offset=3 * n, use_random_inputs=not philox, use_step_counter=True)
    for t in range(steps):
        st = make_state(spec, n, seed=100 + t)
        if t > 0:   # fresh physics + a fresh policy action every step; the manager state chains from the previous step
            st = {k: v for k, v in st.items() if k in nat._STATE_FIELDS or k in ("new_action", "cmd_uniforms", "obs_uniforms_policy", "obs_uniforms_critic")}
        for eng, b in ((eng_a, ba), (eng_b, bb)):
            b.load_logical(st)
            if philox:
                b.cmd_uniforms, b.obs_uniforms = None, [None, None]
            eng.process_action(b)
            eng.step_pre_reset(b, **rng)
            eng.step_post_reset(b, **rng)
        torch.cuda.synchronize()
        _assert_identical(_snapshot(ba), _snapshot(bb), f"{key} N={n} cfg={v2_cfg} philox={philox} step {t}")
        if t == 0 and key.endswith("rough") or t == 0 and key == "go2_rough":
            assert int(bb.n_reset.item()) > 0, "the synthetic state must exercise the reset path"
    assert eng_b.cluster_config(n)["launches"] == 2 * steps, "the cluster kernels did not run"
    assert eng_a.cluster_config(n)["launches"] == 0
    eng_a.close()
    eng_b.close()


@pytest.mark.parametrize("v2_cfg", CONFIGS)
@pytest.mark.parametrize("philox", [False, True])
def test_go2_rough_bit_identical(native_lib, v2_cfg, philox):
    _three_steps("go2_rough", 4096, v2_cfg, philox)


@pytest.mark.parametrize("key", ["a1_flat", "go2_flat", "g1_rough", "g1_rough_37", "g1_flat", "a1_rough"])
@pytest.mark.parametrize("v2_cfg", CONFIGS)
def test_every_baked_task_bit_identical(native_lib, key, v2_cfg):
    _three_steps(key, 1024, v2_cfg, philox=True, steps=2)


@pytest.mark.parametrize("cfg", ["4x4x16", "1x1x8"])
@pytest.mark.parametrize("n", [128, 256, 640, 16384])
def test_env_counts(native_lib, n, cfg):
    """One cluster, an odd number of clusters, several waves of clusters."""
    _three_steps("go2_rough", n, cfg, philox=True, steps=2)


def test_default_config_choice_and_fallback(native_lib):
    """Without RL_MDPSTEP_V2_CFG: one tile per CTA, 16 warps up to a few waves of tiles, 8 warps beyond; env counts that are
    not a multiple of 32 and IsaacLab-shaped tensors run the general kernel."""
    from robot_lab_b200.engine import MdpStepEngine

    cfg, spec = H.make_spec("go2_rough")
    eng = MdpStepEngine(spec, "cuda:0")
    cc = eng.cluster_config(4096)
    assert (cc["cluster_size"], cc["tiles_per_cta"], cc["warps_per_cta"]) == (1, 1, 16)
    cc = eng.cluster_config(65536)
    assert (cc["cluster_size"], cc["tiles_per_cta"], cc["warps_per_cta"]) == (1, 1, 8)
    assert eng.cluster_config(4096 + 32)["cluster_size"] == 1
    assert eng.cluster_config(1000)["cluster_size"] == 0
    for n, layout, want in ((1000, "soa", 0), (1024, "aos", 0), (1024, "soa", 2)):
        st = make_state(spec, n)
        b = eng.new_buffers(n, layout=layout)
        b.load_logical(st)
        before = eng.cluster_config(n)["launches"]
        eng.step_pre_reset(b)
        eng.step_post_reset(b)
        torch.cuda.synchronize()
        assert eng.cluster_config(n)["launches"] - before == want
    eng.close()


@pytest.mark.parametrize("key", ["go2_rough", "g1_rough"])
@pytest.mark.parametrize("v2_cfg", CONFIGS)
def test_v2_two_launch_step_matches_oracle(native_lib, monkeypatch, key, v2_cfg):
    """The env step in the reference's order through the cluster kernels against the CPU oracle (noise as input): the
    body of test_gpu_step_parity.test_two_launch_step_in_reference_order with an engine pinned to one configuration."""
    import test_gpu_step_parity as T

    engines = []

    def pinned(spec):
        engines.append(_engine(spec, v2_cfg))
        return engines[-1]

    monkeypatch.setattr(T, "_engine", pinned)
    launches = []
    from robot_lab_b200.engine import MdpStepEngine

    orig_close = MdpStepEngine.close

    def counting_close(self):
        launches.append(self.cluster_config(2048)["launches"])
        orig_close(self)

    monkeypatch.setattr("robot_lab_b200.engine.MdpStepEngine.close", counting_close)
    T.test_two_launch_step_in_reference_order(native_lib, key, 2048)
    assert launches and launches[-1] == 2, "the cluster kernels did not run"


@pytest.mark.parametrize("v2_cfg", ["1x1x16", "1x1x8", "4x4x16"])
def test_reset_ids_by_lookback_under_graph_replay(native_lib, v2_cfg):
    """The pre-reset kernel orders the reset ids by a decoupled look-back whose status words carry an epoch that lives in
    device memory: a captured graph replays the SAME kernel arguments, so every replay has to see a fresh epoch; engines
    are reused across env counts (stale status words of a larger launch), and one tile / many windows of tiles are the
    edge cases of the look-back. reset_ids == nonzero(done) ascending, n_reset == count, every time."""
    cfg, spec = H.make_spec("go2_rough")
    eng = _engine(spec, v2_cfg)
    gmul = int(v2_cfg.split("x")[1])
    rng = dict(seed=3, use_random_inputs=False, use_step_counter=True)
    for n in (32 * gmul * 40, 32 * gmul, 32 * gmul * 129, 32 * gmul * 3):   # big, one tile (cluster), > 4 windows, small again
        bufs = []
        for i in range(3):
            b = eng.new_buffers(n)
            st = make_state(spec, n, seed=50 + i)
            b.load_logical(st)
            b.cmd_uniforms, b.obs_uniforms = None, [None, None]
            bufs.append(b)
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            for b in bufs:                                   # scratch allocation outside the capture
                eng.step_pre_reset(b, **rng)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for b in bufs:
                    eng.step_pre_reset(b, **rng)
            for rep in range(3):
                for b in bufs:                               # poison the outputs: a replay has to rewrite them
                    b.reset_ids.fill_(-7)
                    b.n_reset.fill_(-1)
                g.replay()
                stream.synchronize()
                for i, b in enumerate(bufs):
                    done = (b.terminated.bool() | b.truncated.bool()).nonzero().flatten()
                    k = int(b.n_reset.item())
                    assert k == done.numel(), f"N={n} cfg={v2_cfg} replay {rep} set {i}: n_reset {k} != {done.numel()}"
                    assert torch.equal(b.reset_ids[:k].long(), done), f"N={n} cfg={v2_cfg} replay {rep} set {i}: ids differ"
        assert eng.cluster_config(n)["launches"] > 0, "the cluster kernels did not run"
    eng.close()

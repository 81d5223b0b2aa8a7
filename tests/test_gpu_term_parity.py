"""Per-term GPU parity: every reward term evaluated alone through ``rl_term_eval`` (the term-function protocol,
V/mdp/rewards.py:22 ff.) against the oracle's restatement of that term."""

import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200.synthetic import make_state

pytestmark = pytest.mark.gpu


def _check_terms(spec, terms, n=2048):
    from robot_lab_b200.engine import MdpStepEngine

    st = make_state(spec, n)
    st["terminated"] = torch.rand(n, generator=torch.Generator().manual_seed(3)) < 0.1
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n)
    b.load_logical(st)
    term_dev = st["terminated"].to(torch.uint8).cuda()
    d = port.Derived(st, spec)
    failures = []
    for t in terms:
        got = eng.term_eval(t, b, terminated=term_dev).cpu()
        ref = port.reward_term(t, st, spec, d)
        # a term that subtracts an O(100 N) threshold from an fp32 norm carries one ulp of THAT scale (7.6e-6)
        atol = 1e-6 * max(1.0, abs(t.p[0])) if t.type_name == "contact_forces" else H.ATOL
        try:
            torch.testing.assert_close(got, ref, rtol=H.RTOL, atol=atol)
        except AssertionError as e:
            failures.append(f"{t.name}: {str(e).splitlines()[-3:]}")
    eng.close()
    assert not failures, "\n".join(failures)


@pytest.mark.parametrize("key", ["a1_flat", "go2_rough", "g1_rough", "g1_rough_37"])
def test_active_terms(native_lib, key):
    cfg, spec = H.make_spec(key)
    _check_terms(spec, spec.rewards)


def test_whole_catalogue_including_inactive_terms(native_lib):
    """The terms the in-scope tasks leave at weight 0 (V/velocity_env_cfg.py:379-644) are implemented too."""
    cfg, spec = H.make_catalogue_spec()
    assert spec.K >= 34
    _check_terms(spec, spec.rewards)

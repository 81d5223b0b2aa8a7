"""ContactSensor update [IL] through the C-ABI against the oracle restatement."""
import pytest
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200.engine import MdpStepEngine
from robot_lab_b200.synthetic import make_state

pytestmark = pytest.mark.gpu


def _forces(n, B, seed):
    g = torch.Generator().manual_seed(seed)
    mag = -30.0 * torch.log1p(-torch.rand(n, B, generator=g))
    mag = torch.where(torch.rand(n, B, generator=g) < 0.4, mag, torch.zeros(()))
    d = torch.randn(n, B, 3, generator=g)
    f = d / d.norm(dim=-1, keepdim=True).clamp(min=1e-9) * mag.unsqueeze(-1)
    f[0] = 0.0
    f[0, :, 0] = 1.0          # exactly on the threshold: |F| == 1.0 is NOT a contact (strict >)
    return f


@pytest.mark.parametrize("key,n,layout", [("go2_rough", 1000, "soa"), ("g1_rough", 257, "aos"), ("a1_flat", 64, "soa")])
def test_sub_steps_match_oracle(native_lib, key, n, layout):
    cfg, spec = H.make_spec(key)
    st = make_state(spec, n)
    eng = MdpStepEngine(spec, "cuda:0")
    b = eng.new_buffers(n, layout=layout)
    b.load_logical(st)
    ref = {k: st[k].clone() for k in ("net_forces_w_history", "current_air_time", "last_air_time", "current_contact_time",
                                      "last_contact_time")}
    dt = 0.005
    for sub in range(6):      # more sub-steps than the history is long
        f = _forces(n, spec.B, 100 + sub)
        eng.contact_sensor_update(b, f.cuda(), dt)
        ref.update(port.contact_sensor_update(spec, ref, f, dt))
    torch.cuda.synchronize()
    for k, want in ref.items():
        got = b.logical(k).cpu().contiguous()
        torch.testing.assert_close(got, want, rtol=0, atol=0, msg=k)   # copies, adds of dt and selects: bit-exact
    eng.close()


def test_ring_slot_mode_feeds_the_same_terms(native_lib):
    """Writing only slot (sub mod T) moves no data; every history consumer on the path takes max over the history
    axis, so the terms see the same values as with the rolled tensor."""
    cfg, spec = H.make_spec("go2_rough")
    n = 512
    st = make_state(spec, n)
    eng = MdpStepEngine(spec, "cuda:0")
    roll, ring = eng.new_buffers(n), eng.new_buffers(n)
    roll.load_logical(st)
    ring.load_logical(st)
    for sub in range(7):
        f = _forces(n, spec.B, 7 + sub).cuda()
        eng.contact_sensor_update(roll, f, 0.005)
        eng.contact_sensor_update(ring, f, 0.005, ring_slot=sub % spec.T)
    torch.cuda.synchronize()
    h_roll, h_ring = roll.logical("net_forces_w_history"), ring.logical("net_forces_w_history")
    torch.testing.assert_close(h_roll.norm(dim=-1).max(dim=1).values, h_ring.norm(dim=-1).max(dim=1).values, rtol=0, atol=0)
    for k in ("current_air_time", "last_air_time", "current_contact_time", "last_contact_time"):
        assert torch.equal(roll.logical(k), ring.logical(k)), k
    eng.step(roll)
    eng.step(ring)
    torch.cuda.synchronize()
    assert torch.equal(roll.reward, ring.reward)
    assert torch.equal(roll.terminated, ring.terminated)
    eng.close()

"""Hand-checked sequence for the ContactSensor restatement (oracle/mdp_port.contact_sensor_update)."""
import torch

import helpers as H
from oracle import mdp_port as port
from robot_lab_b200.synthetic import make_state


def test_air_time_state_machine_on_a_known_sequence():
    cfg, spec = H.make_spec("go2_rough")
    st = make_state(spec, 2)
    names_h = list(spec.layout.hist_body_names)
    foot = names_h.index(spec.layout.time_body_names[0])
    s = {"net_forces_w_history": torch.zeros(2, spec.T, spec.B, 3),
         "current_air_time": torch.zeros(2, spec.Bt), "last_air_time": torch.zeros(2, spec.Bt),
         "current_contact_time": torch.zeros(2, spec.Bt), "last_contact_time": torch.zeros(2, spec.Bt)}
    dt = 0.005

    def sub(force):
        f = torch.zeros(2, spec.B, 3)
        f[0, foot, 2] = force
        s.update(port.contact_sensor_update(spec, s, f, dt))

    for _ in range(3):
        sub(0.0)                                   # in the air: 3 sub-steps
    assert torch.allclose(s["current_air_time"][0, 0], torch.tensor(3 * dt))
    sub(50.0)                                      # touch-down: last_air = 3 dt + dt (the elapsed sub-step counts)
    assert torch.allclose(s["last_air_time"][0, 0], torch.tensor(4 * dt))
    assert s["current_air_time"][0, 0] == 0 and torch.allclose(s["current_contact_time"][0, 0], torch.tensor(dt))
    sub(1.0)                                       # |F| == threshold is not a contact: lift-off
    assert torch.allclose(s["last_contact_time"][0, 0], torch.tensor(2 * dt))
    assert s["current_contact_time"][0, 0] == 0 and torch.allclose(s["current_air_time"][0, 0], torch.tensor(dt))
    # history: newest first, three samples kept
    assert s["net_forces_w_history"][0, 0, foot, 2] == 1.0 and s["net_forces_w_history"][0, 1, foot, 2] == 50.0
    assert s["net_forces_w_history"][0, 2, foot, 2] == 0.0

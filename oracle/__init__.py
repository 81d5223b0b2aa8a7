"""CPU oracle of the MDP step - TEST INFRASTRUCTURE ONLY (see oracle/mdp_port.py for the scope and the
parity-pinning status). Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs - never from robot_lab_b200."""

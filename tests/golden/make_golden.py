"""Generate the committed golden fixtures by running the REFERENCE's own functions (unmodified files under
/root/reference, loaded through oracle/isaaclab_shim.py) on seeded synthetic state.

    python tests/golden/make_golden.py          # build container only (needs /root/reference)

Each ``<task>.npz`` holds: the synthetic input state (logical shapes, fp32), the raw value of every reward term
the reference owns (V/mdp/rewards.py functions incl. the GaitReward class), the reference's
UniformThresholdVelocityCommand.compute() result (V/mdp/commands.py over a restated [IL] base class) and the two
reference observation functions (V/mdp/observations.py). IsaacLab-owned terms have no reference source
("parity unpinned", oracle/mdp_port.py) and are NOT in the fixtures.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import helpers as H  # noqa: E402
from oracle import isaaclab_shim, ref_harness  # noqa: E402
from robot_lab_b200.synthetic import make_state  # noqa: E402

N = {"a1_flat": 64, "go2_flat": 48, "go2_rough": 48, "g1_rough": 48, "g1_rough_37": 32}
SEED = 20260922


def main():
    assert isaaclab_shim.reference_available(), "needs /root/reference"
    out_dir = Path(__file__).resolve().parent
    for key, n in list(N.items()) + [("go2_catalogue", 48)]:
        cfg, spec = H.make_catalogue_spec() if key == "go2_catalogue" else H.make_spec(key)
        st = make_state(spec, n, seed=SEED)
        arrays = {f"in/{k}": v.numpy() for k, v in st.items()}
        for name, val in ref_harness.reference_reward_terms(cfg, spec, st).items():
            if val is not None:
                arrays[f"reward/{name}"] = val.float().numpy()
        ter = "plane" if "flat" in key else "generator"
        if key == "go2_catalogue":
            np.savez_compressed(out_dir / f"{key}.npz", **arrays)
            print(key, n, "envs,", sum(1 for k in arrays if k.startswith("reward/")), "reference-owned reward terms")
            continue
        for name, val in ref_harness.reference_command_compute(spec, st, st["cmd_uniforms"], ter).items():
            arrays[f"command/{name}"] = val.numpy()
        obs = ref_harness.reference_observation_terms(spec, st)
        arrays["obs/phase"] = obs["phase"].numpy()
        arrays["obs/joint_pos_rel_without_wheel"] = obs["joint_pos_rel_without_wheel"].numpy()
        arrays["obs/_wheel_ids"] = obs["_wheel_ids"].numpy()
        np.savez_compressed(out_dir / f"{key}.npz", **arrays)
        print(key, n, "envs,", sum(1 for k in arrays if k.startswith("reward/")), "reference-owned reward terms")


def reset_event():
    """reset_state_<task>.npz: inputs (ids, uniforms, env origins) and what the reference's reset_root_state_uniform
    (V/mdp/events.py:205-271) writes into the simulator for them (Go2 rough ranges, GO2/rough_env_cfg.py:56-73)."""
    from robot_lab_b200.cfg import ResetStateCfg

    out_dir = Path(__file__).resolve().parent
    for key, n in (("go2_rough", 96), ("g1_rough", 40)):
        cfg, spec = H.make_spec(key)
        st = make_state(spec, n, seed=SEED)
        g = torch.Generator().manual_seed(SEED + 1)
        uniforms = torch.rand(12 + 2 * spec.J, n, generator=g)
        origins = torch.randn(n, 3, generator=g) * 20.0
        ids = torch.randperm(n, generator=g)[: n // 3].sort().values.int()
        ref = ref_harness.reference_reset_root_state(spec, st, ids, ResetStateCfg.go2_rough(), origins, uniforms)
        arrays = {"ids": ids.numpy(), "uniforms": uniforms.numpy(), "env_origins": origins.numpy()}
        arrays.update({f"out/{k}": v.numpy() for k, v in ref.items()})
        np.savez_compressed(out_dir / f"reset_state_{key}.npz", **arrays)
        print("reset_state", key, n, "envs,", len(ids), "reset")


if __name__ == "__main__":
    if "--reset-event-only" not in sys.argv:
        main()
    reset_event()

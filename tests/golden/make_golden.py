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
            if name != "was_on_pit":   # all False without a "pits" sub-terrain; pit_terrain_*.npz covers that branch
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


PIT_SUB_TERRAINS = ("pyramid_stairs", "pits", "boxes", "random_rough", "hf_pyramid_slope")
PIT_PROPORTIONS = (0.2, 0.15, 0.25, 0.3, 0.1)


def pits():
    """pit_terrain_<task>.npz: a generator terrain WITH a "pits" sub-terrain (none of the reference's task configs
    has one, so the golden vectors of main() never reach that code): the reference's is_robot_on_terrain /
    is_env_assigned_to_terrain (V/mdp/utils.py:44-127), UniformThresholdVelocityCommand.compute() with the pit
    branch of _update_command live (V/mdp/commands.py:61-85) and the pit branch of reset_root_state_uniform
    (V/mdp/events.py:232-244)."""
    from robot_lab_b200 import terrain as terrain_host
    from robot_lab_b200.cfg import ResetStateCfg, TerrainCfg

    out_dir = Path(__file__).resolve().parent
    ter_cfg = TerrainCfg(sub_terrains=PIT_SUB_TERRAINS, proportions=PIT_PROPORTIONS)
    for key, n in (("go2_rough", 512), ("g1_rough", 256)):
        cfg, spec = H.make_spec(key)
        st = make_state(spec, n, seed=SEED + 7)
        g = torch.Generator().manual_seed(SEED + 8)
        st["root_pos_w"] = torch.stack([(torch.rand(n, generator=g) - 0.5) * 70.0, (torch.rand(n, generator=g) - 0.5) * 150.0,
                                        torch.rand(n, generator=g)], dim=1)
        origins = terrain_host.grid_origins(ter_cfg)
        origins[:, :, 2] = torch.rand(origins.shape[:2], generator=g)
        types = torch.randint(0, ter_cfg.num_cols, (n,), generator=g)
        was = torch.rand(n, generator=g) < 0.35
        ter = ref_harness.fake_terrain(ter_cfg, "generator", n, origins, types)
        keep = ("root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w", "command", "heading_target", "time_left",
                "is_heading_env", "is_standing_env", "metric_error_vel_xy", "metric_error_vel_yaw", "cmd_uniforms")
        arrays = {f"in/{k}": st[k].numpy() for k in keep}   # everything the command term and the reset event read
        arrays.update({"terrain_origins": origins.numpy(), "terrain_types": types.numpy(), "was_on_pit": was.numpy()})
        ref = ref_harness.reference_command_compute(spec, st, st["cmd_uniforms"], "generator", terrain=ter, was_on_pit=was)
        arrays.update({f"command/{k}": v.numpy() for k, v in ref.items()})
        utils = ref_harness.reference_terrain_utils()
        env = ref_harness.FakeEnv(spec, st)
        env.scene.terrain = ter
        for name in ("pits", "boxes"):
            arrays[f"on_terrain/{name}"] = utils.is_robot_on_terrain(env, name).numpy()
            arrays[f"assigned/{name}"] = utils.is_env_assigned_to_terrain(env, name).numpy()
        uniforms = torch.rand(12 + 2 * spec.J, n, generator=g)
        env_origins = torch.randn(n, 3, generator=g) * 20.0
        ids = torch.randperm(n, generator=g)[: n // 3].sort().values.int()
        rr = ref_harness.reference_reset_root_state(spec, st, ids, ResetStateCfg.go2_rough(), env_origins, uniforms, terrain=ter)
        arrays.update({"reset/ids": ids.numpy(), "reset/uniforms": uniforms.numpy(), "reset/env_origins": env_origins.numpy()})
        arrays.update({f"reset/out/{k}": v.numpy() for k, v in rr.items()})
        np.savez_compressed(out_dir / f"pit_terrain_{key}.npz", **arrays)
        print("pit_terrain", key, n, "envs,", int(ref["was_on_pit"].sum()), "on pits,", int((was & ~ref["was_on_pit"]).sum()), "left a pit")


if __name__ == "__main__":
    if "--pits-only" in sys.argv:
        pits()
    else:
        if "--reset-event-only" not in sys.argv:
            main()
        reset_event()
        pits()

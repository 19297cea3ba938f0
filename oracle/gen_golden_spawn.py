"""TEST INFRASTRUCTURE -- spawn distributions of the REFERENCE's own reset (crowd_sim_dict.py:105-203 ->
generate_robot_humans, crowd_sim.py:555-663, 359-393, 296-357) for the distributional parity test of the device reset
(SURVEY.md test plan T5; DESIGN.md deviations D1 RNG / D2 bounded rejection).

Build container only (imports /root/reference under oracle/ref_import.py shims):

    python -m oracle.gen_golden_spawn            # writes tests/golden/spawn_<case>.npz

Per case the reference env is reset `RESETS` times in the train phase (every reset advances its case counter, so
every episode has its own seed) and the spawned state is reduced to per-feature samples; the fixture stores 257
quantiles of every feature (the empirical CDF the CUDA reset is compared with) plus the constraint-violation counts.
"""
import os
import sys

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
RESETS = 3000
SCENARIOS = ["circle_crossing", "square_crossing", "parallel_traffic", "perpendicular_traffic"]
CASES = {
    "circle_h5": dict(over={"sim.train_val_sim": ["circle_crossing"]}),
    "square_h5": dict(over={"sim.train_val_sim": ["square_crossing"]}),
    "parallel_h5": dict(over={"sim.train_val_sim": ["parallel_traffic"]}),
    "perpendicular_h5": dict(over={"sim.train_val_sim": ["perpendicular_traffic"]}),
    "circle_h10_unicycle": dict(over={"sim.train_val_sim": ["circle_crossing"], "sim.human_num": 10,
                                      "action_space.kinematics": "unicycle", "env.time_step": 0.1}),
    "mixed_h5": dict(over={}),          # the default four scenarios, chosen at random per episode: scenario frequencies
    # the group environment (circle groups of static humans, crowd_sim.py:476-622)
    "group_h8": dict(over={"sim.group_human": True, "sim.human_num": 8, "sim.train_val_sim": ["circle_crossing"]}),
    "group_h16_mixed": dict(over={"sim.group_human": True, "sim.human_num": 16}),
}


def features(robot, humans, discomfort=0.25):
    """robot [9] and humans [H, 9] in Agent.get_full_state_list order (px, py, vx, vy, radius, gx, gy, v_pref, theta)
    -> dict of 1-D sample arrays.  Shared by the generator and tests/test_gpu_spawn.py."""
    hp, hg, hr, hv = humans[:, 0:2], humans[:, 5:7], humans[:, 4], humans[:, 7]
    rp, rg, rr = robot[0:2], robot[5:7], robot[4]
    H = humans.shape[0]
    d = np.linalg.norm(hp[:, None, :] - hp[None, :, :], axis=-1) - hr[:, None] - hr[None, :]
    clear_hh = d[np.triu_indices(H, 1)] if H > 1 else np.zeros(0)
    clear_rh = np.linalg.norm(hp - rp, axis=-1) - hr - rr
    return {
        "human_px": hp[:, 0], "human_py": hp[:, 1], "human_gx": hg[:, 0], "human_gy": hg[:, 1],
        "human_radial": np.linalg.norm(hp, axis=-1), "human_goal_dist": np.linalg.norm(hg - hp, axis=-1),
        "human_radius": hr, "human_v_pref": hv,
        "robot_px": rp[0:1], "robot_py": rp[1:2], "robot_gx": rg[0:1], "robot_gy": rg[1:2],
        "robot_goal_dist": np.array([np.linalg.norm(rg - rp)]),
        "min_clear_human_human": np.array([clear_hh.min()]) if clear_hh.size else np.zeros(0),
        "min_clear_robot_human": np.array([clear_rh.min()]),
    }


def run_case(name):
    from . import ref_harness

    case = CASES[name]
    cfg = ref_harness.make_reference_config(**dict(case["over"], **{"training.num_processes": 16, "env.seed": 0}))
    renv = ref_harness.RefEnv(cfg, n_envs=16, phase="train")
    env = renv.env
    samples, scen = {}, []
    for _ in range(RESETS):
        env.reset()
        robot = np.array(env.robot.get_full_state_list(), dtype=np.float64)
        humans = np.array([h.get_full_state_list() for h in env.humans], dtype=np.float64)
        for k, v in features(robot, humans).items():
            samples.setdefault(k, []).append(v)
        scen.append(SCENARIOS.index(env.current_scenario))
    q = np.linspace(0.0, 1.0, 257)
    out = {"quantile_grid": q, "resets": np.array(RESETS), "human_num": np.array(cfg.sim.human_num),
           "scenario_counts": np.bincount(np.array(scen), minlength=4)}
    for k, parts in samples.items():
        x = np.concatenate(parts)
        out["q_" + k] = np.quantile(x, q)
        out["n_" + k] = np.array(x.size)
    out["overrides_keys"] = np.array(list(case["over"].keys()))
    out["overrides_vals"] = np.array([repr(v) for v in case["over"].values()])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "spawn_%s.npz" % name), **out)
    print(name, "resets", RESETS, "min clearances", out["q_min_clear_human_human"][0] if "q_min_clear_human_human" in out else None,
          out["q_min_clear_robot_human"][0], "scenarios", out["scenario_counts"])


if __name__ == "__main__":
    from . import crowd_oracle

    crowd_oracle.build()
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name)

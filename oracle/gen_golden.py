"""TEST INFRASTRUCTURE -- generates tests/golden/step_*.npz by executing the
REFERENCE's own CrowdSimDict.step (imported from /root/reference under
oracle/ref_import.py) on seeded injected states, and prints how the C oracle
(oracle/crowd_oracle.c) compares.  Run in the build container only:

    python -m oracle.gen_golden            # writes tests/golden/step_<case>.npz

Each fixture holds the injected inputs (float32), the reference's outputs
(float64 as the reference computed them) and the JSON of the config overrides
so the tests can rebuild the same CnConfig without the reference.
Goal re-sampling is RNG-driven (global MT19937 in the reference, Philox here),
so the fixtures run with humans.random_goal_changing = end_goal_changing =
False; which humans WOULD trigger an end-goal change is recorded separately.
"""
import json
import os
import sys

import numpy as np

from crowdnav_dsrnn_b200 import abi
from . import crowd_oracle, ref_harness, state_sampler

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# name -> (reference Config overrides, n_envs, seed)
CASES = {
    # BASELINE.json configs[0]: default holonomic, 5 humans, circle crossing
    "c1_holonomic_h5": ({"sim.train_val_sim": ["circle_crossing"]}, 96, 101),
    # configs[1]: unicycle, 10 humans, mixed scenarios, dt = 0.1 as data/example_model_unicycle/configs/config.py:20
    "c2_unicycle_h10": ({"action_space.kinematics": "unicycle", "sim.human_num": 10, "env.time_step": 0.1,
                         "reward.discomfort_penalty_factor": 10 * 0.1}, 96, 202),
    # configs[2]: 20 humans, robot FOV 90 degrees
    "c3_fov_h20": ({"sim.human_num": 20, "robot.FOV": 0.5}, 64, 303),
    # configs[3]: social metrics geometry (circle_radius 4)
    "c4_social_h5": ({"test.social_metrics": True, "sim.circle_radius": 4, "env.test_size": 2000}, 64, 404),
    # side preference: one human, goal changing off by construction
    "c4_sidepref_h1": ({"test.side_preference": True, "sim.test_sim": ["side_pref_passing"],
                        "sim.train_val_sim": ["side_pref_passing"], "sim.circle_radius": 4, "sim.human_num": 1,
                        "env.test_size": 200}, 48, 505),
    # robot visible to the humans' ORCA, unicycle with dt 0.25, exponential reward off
    "x_robot_visible_h5": ({"robot.visible": True}, 48, 606),
    # limited HUMAN field of view (dummy-human substitution, crowd_sim.py:1139-1142)
    "x_human_fov_h6": ({"humans.FOV": 1.0, "sim.human_num": 6}, 48, 707),
    "x_unicycle_fov_h5": ({"action_space.kinematics": "unicycle", "robot.FOV": 1.0}, 48, 808),
    # social-force humans (crowd_nav/policy/social_force.py), plain and with limited human FOV + visible robot
    "n4_social_force_h5": ({"humans.policy": "social_force"}, 96, 909),
    "n4_social_force_fov_h8": ({"humans.policy": "social_force", "humans.FOV": 1.0, "robot.visible": True,
                                "sim.human_num": 8}, 48, 1010),
    "n4_social_force_uni_h10": ({"humans.policy": "social_force", "action_space.kinematics": "unicycle", "sim.human_num": 10,
                                 "env.time_step": 0.1, "reward.discomfort_penalty_factor": 10 * 0.1}, 48, 1111),
}
COMMON = {"humans.random_goal_changing": False, "humans.end_goal_changing": False}


def flat_cfg(ref_cfg, n_envs):
    return abi.flatten_config(ref_cfg, n_envs, phase="train")


def reference_vs_oracle(over, n_envs, seed):
    """Run the reference's own step and the C oracle on the same seeded injected states; returns (inputs, reference outputs,
    comparison report).  Used by run_case (fixtures) and by tests/test_oracle_live_reference.py (fresh seeds)."""
    over = dict(COMMON, **over)
    ref_cfg = ref_harness.make_reference_config(**over)
    cfg = flat_cfg(ref_cfg, n_envs)
    H = cfg.human_num
    inp = state_sampler.sample(cfg, n_envs, seed)

    # ---- reference
    ref = {k: [] for k in ("robot_node", "temporal_edges", "spatial_edges", "visible", "reward", "done", "event",
                           "dmin", "aggregate_nav_time", "path_violation", "personal_violation", "jerk_cost",
                           "dist_to_goal", "speed_violation", "scenario", "side_left", "side_right", "separation",
                           "robot", "humans", "belief", "extras", "global_time", "end_goal_trigger")}
    renv = ref_harness.RefEnv(ref_cfg, n_envs=n_envs)
    for e in range(n_envs):
        renv.inject(inp["robot"][e], inp["humans"][e], inp["belief"][e], inp["extras"][e], inp["counters"][e])
        o = renv.step(inp["action"][e])
        st = renv.extract()
        for k in ("side_left", "side_right", "separation"):
            o.setdefault(k, 0.0)
        for k, v in o.items():
            ref[k].append(v)
        for k, v in st.items():
            ref[k].append(v)
        trig = [bool(np.linalg.norm((h.gx - h.px, h.gy - h.py)) < h.radius) for h in renv.env.humans]
        ref["end_goal_trigger"].append(trig)
    ref = {k: np.array(v) for k, v in ref.items()}

    # ---- C oracle on the same inputs
    st = crowd_oracle.OracleState(n_envs, H)
    for f in ("robot", "humans", "belief", "extras", "counters"):
        getattr(st, f)[...] = inp[f]
    out = crowd_oracle.step(cfg, st, inp["action"], auto_reset=False)

    return over, inp, ref, compare(ref, out, st, H)


def run_case(name, over, n_envs, seed):
    over, inp, ref, rep = reference_vs_oracle(over, n_envs, seed)
    print(f"[{name}] N={n_envs}: " + ", ".join(f"{k}={v}" for k, v in rep.items()))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, f"step_{name}.npz"),
        overrides=json.dumps(over), n_envs=n_envs, seed=seed,
        **{"in_" + k: v for k, v in inp.items()}, **{"ref_" + k: v for k, v in ref.items()})
    return rep


def compare(ref, out, st, H):
    """max abs error / mismatch counts of the C oracle against the reference outputs."""
    vis = np.array([[(int(m) >> i) & 1 for i in range(H)] for m in out.visible_mask], bool)
    rep = {
        "flag_mismatch": int((out.done.astype(bool) != ref["done"]).sum() + (out.event != ref["event"]).sum()),
        "vis_mismatch": int((vis != ref["visible"]).sum()),
        "int_info_mismatch": int(sum((out.info[:, abi.INFO_COLUMNS[k]] != ref[k]).sum() for k in
                                     ("aggregate_nav_time", "path_violation", "personal_violation",
                                      "speed_violation", "side_left", "side_right"))),
        "reward": float(np.abs(out.reward - ref["reward"]).max()),
        "robot_pv": float(np.abs(st.robot[:, [0, 1, 2, 3, 8]] - ref["robot"][:, [0, 1, 2, 3, 8]]).max()),
        "human_pv": float(np.abs(st.humans[:, :, 0:4] - ref["humans"][:, :, 0:4]).max()),
        "belief": float(np.abs(st.belief - ref["belief"]).max()),
        "obs": float(max(np.abs(out.robot_node.reshape(ref["robot_node"].shape) - ref["robot_node"]).max(),
                         np.abs(out.temporal_edges.reshape(ref["temporal_edges"].shape) - ref["temporal_edges"]).max(),
                         np.abs(out.spatial_edges - ref["spatial_edges"]).max())),
        "extras": float(np.abs(st.extras - ref["extras"]).max()),
        "jerk": float(np.abs(out.info[:, abi.INFO_COLUMNS["jerk_cost"]] - ref["jerk_cost"]).max()),
        "dist_to_goal": float(np.abs(out.info[:, abi.INFO_COLUMNS["dist_to_goal"]] - ref["dist_to_goal"]).max()),
    }
    danger = ref["event"] == 1
    if danger.any():
        rep["dmin"] = float(np.abs(out.info[danger, 0] - ref["dmin"][danger]).max())
    return rep


def main():
    crowd_oracle.build()
    names = sys.argv[1:] or list(CASES)
    for name in names:
        over, n, seed = CASES[name]
        run_case(name, over, n, seed)


if __name__ == "__main__":
    main()

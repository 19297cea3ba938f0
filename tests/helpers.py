"""Shared helpers of the parity tests (test infrastructure)."""
import glob
import json
import os

import numpy as np

from crowdnav_dsrnn_b200 import Config, abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEP_CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "step_*.npz")))
DSRNN_CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "dsrnn_*.npz")))

# tolerances stated by BASELINE.json's north_star
TOL_POS = 1e-4      # positions / velocities after one step [m, m/s]
TOL_REWARD = 1e-5
TOL_NET_REL = 1e-3  # DS-RNN action means and values, relative


def config_from_overrides(over):
    """This package's Config with the dotted overrides a golden fixture records."""
    cfg = Config()
    for key, val in over.items():
        sec, _, attr = key.partition(".")
        setattr(getattr(cfg, sec), attr, val)
    return cfg


def load_step_case(name):
    d = np.load(os.path.join(GOLDEN, "step_%s.npz" % name))
    over = json.loads(str(d["overrides"]))
    cfg_obj = config_from_overrides(over)
    n = int(d["n_envs"])
    return d, cfg_obj, abi.flatten_config(cfg_obj, n, phase="train"), n


def visible_bits(mask, H):
    mask = np.asarray(mask).astype(np.int64) & 0xFFFFFFFF
    return ((mask[:, None] >> np.arange(H)[None, :]) & 1).astype(bool)


def check_against_reference(d, cfg, out, state, where=""):
    """Assert `out`/`state` (dicts of numpy arrays in the canonical layout) against a golden fixture `d`
    with the north_star's tolerances: flags bit-exact, dynamics 1e-4, reward 1e-5."""
    H = cfg.human_num
    col = abi.INFO_COLUMNS
    assert np.array_equal(out["done"].astype(bool), d["ref_done"]), where + " done flags"
    assert np.array_equal(out["event"], d["ref_event"]), where + " event (collision/success/timeout) flags"
    assert np.array_equal(visible_bits(out["visible_mask"], H), d["ref_visible"]), where + " FOV visibility mask"
    for k in ("aggregate_nav_time", "path_violation", "personal_violation", "speed_violation", "side_left", "side_right"):
        assert np.array_equal(out["info"][:, col[k]], d["ref_" + k].astype(np.float32)), where + " info " + k
    assert np.abs(out["reward"] - d["ref_reward"]).max() <= TOL_REWARD, where + " reward"
    assert np.abs(state["robot"][:, [0, 1, 2, 3, 8]] - d["ref_robot"][:, [0, 1, 2, 3, 8]]).max() <= TOL_POS
    assert np.abs(state["humans"][:, :, 0:4] - d["ref_humans"][:, :, 0:4]).max() <= TOL_POS, where + " human p,v"
    assert np.abs(state["belief"] - d["ref_belief"]).max() <= TOL_POS, where + " belief"
    assert np.abs(state["extras"] - d["ref_extras"]).max() <= TOL_POS, where + " desiredVelocity/potential/last_acc"
    n = len(d["ref_done"])
    assert np.abs(out["robot_node"].reshape(n, 7) - d["ref_robot_node"].reshape(n, 7)).max() <= TOL_POS
    assert np.abs(out["temporal_edges"].reshape(n, 2) - d["ref_temporal_edges"].reshape(n, 2)).max() <= TOL_POS
    assert np.abs(out["spatial_edges"] - d["ref_spatial_edges"]).max() <= TOL_POS, where + " spatial_edges"
    danger = d["ref_event"] == abi.EV_DANGER
    if danger.any():
        assert np.abs(out["info"][danger, col["dmin"]] - d["ref_dmin"][danger]).max() <= TOL_POS
    assert np.abs(out["info"][:, col["jerk_cost"]] - d["ref_jerk_cost"]).max() <= TOL_POS
    assert np.abs(out["info"][:, col["dist_to_goal"]] - d["ref_dist_to_goal"]).max() <= TOL_POS
    # the end-goal trigger |g - p| < r is an index/flag output: bit-exact
    trig = np.linalg.norm(state["humans"][:, :, 5:7].astype(np.float64) - state["humans"][:, :, 0:2], axis=-1) < state["humans"][:, :, 4]
    assert np.array_equal(trig, d["ref_end_goal_trigger"]), where + " end-goal trigger"

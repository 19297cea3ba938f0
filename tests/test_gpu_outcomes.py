"""GPU, distributional parity (BASELINE.json north_star): over 2000-episode test runs the success / collision /
timeout rates of the batched CUDA path must fall within binomial noise of the reference's.  The reference numbers
(tests/golden/outcomes_*.json) come from executing the reference's own env + Policy for 2000 episodes in the build
container (oracle/gen_golden_outcomes.py).  Reset and goal re-sampling use a different RNG and bounded rejection
loops, so this is the test that pins them -- statistically -- against the reference."""
import glob
import json
import math
import os

import numpy as np
import pytest
import torch

from crowdnav_dsrnn_b200.evaluation import evaluate_batched
from crowdnav_dsrnn_b200.model import Policy
from crowdnav_dsrnn_b200.spaces import crowd_spaces
from helpers import GOLDEN, config_from_overrides

pytestmark = pytest.mark.gpu
CASES = sorted(os.path.basename(p)[9:-5] for p in glob.glob(os.path.join(GOLDEN, "outcomes_*.json")))


def _z(p1, n1, p2, n2):
    p = (p1 * n1 + p2 * n2) / (n1 + n2)
    s = math.sqrt(max(p * (1 - p), 1e-4) * (1.0 / n1 + 1.0 / n2))
    return abs(p1 - p2) / s


@pytest.mark.parametrize("case", CASES)
def test_outcome_rates_within_binomial_noise_of_reference(case):
    ref = json.load(open(os.path.join(GOLDEN, "outcomes_%s.json" % case)))
    cfg = config_from_overrides(ref["overrides"])
    ckpt = "holonomic_27776" if "example_model/" in ref["checkpoint"] else "unicycle_55554"
    w = np.load(os.path.join(GOLDEN, "weights_%s.npz" % ckpt))
    obs, act = crowd_spaces(cfg.sim.human_num)
    policy = Policy(obs.spaces, act, base="srnn", base_kwargs=cfg)
    policy.load_state_dict({k: torch.from_numpy(w[k]) for k in w.files})
    policy = policy.cuda()
    n = 4000          # more episodes than the reference run: the batch costs nothing and tightens our side
    got = evaluate_batched(policy, cfg, "cuda:0", episodes=n)
    assert got["unfinished"] == 0
    print(case, "reference", {k: ref[k] for k in ("success", "collision", "timeout", "mean_steps")},
          "cuda", {k: got[k] for k in ("success", "collision", "timeout", "mean_steps")})
    for k in ("success", "collision", "timeout"):
        assert _z(got[k], n, ref[k], ref["episodes"]) <= 4.0, (k, got[k], ref[k])
    assert abs(got["mean_steps_success"] - ref["mean_steps_success"]) <= 0.08 * ref["mean_steps_success"]
    # social metrics SM1-SM5: per-episode means within 5 standard errors (+5 % relative slack for the heavy-tailed ones)
    for k, r in ref.get("social", {}).items():
        g_mean, g_std = got["mean_%s_per_episode" % k], got["std_%s_per_episode" % k]
        se = math.sqrt(r["std"] ** 2 / ref["episodes"] + g_std ** 2 / n)
        assert abs(g_mean - r["mean"]) <= 5.0 * se + 0.05 * abs(r["mean"]) + 1e-9, (k, g_mean, r)
    if cfg.test.side_preference:   # SM6: which side the robot passes on
        for k in ("side_left_episodes", "side_right_episodes"):
            assert _z(got[k], n, ref[k], ref["episodes"]) <= 4.0, (k, got[k], ref[k])
    for scn, r in ref["per_scenario"].items():
        g = got["per_scenario"][scn]
        assert _z(g["success"], g["episodes"], r["success"], r["episodes"]) <= 4.5, (scn, g, r)

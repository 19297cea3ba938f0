"""GPU: `evaluation.evaluate` -- the reference's evaluate() signature and log lines (evaluation.py:283-330, metrics.py,
helper.py:88-101) on top of the batched evaluator."""
import logging
import os

import numpy as np
import pytest
import torch

from crowdnav_dsrnn_b200 import Config
from crowdnav_dsrnn_b200.envs import make_vec_envs
from crowdnav_dsrnn_b200.evaluation import evaluate, evaluate_batched
from crowdnav_dsrnn_b200.model import Policy
from crowdnav_dsrnn_b200.spaces import crowd_spaces

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__()
        self.lines = []

    def emit(self, record):
        self.lines.append(record.getMessage())


def test_evaluate_has_the_reference_signature_and_log_lines():
    cfg = Config()
    cfg.test.social_metrics = True
    cfg = cfg.derive()
    cfg.env.test_size = 256
    w = np.load(os.path.join(GOLDEN, "weights_holonomic_27776.npz"))
    obs, act = crowd_spaces(cfg.sim.human_num)
    policy = Policy(obs.spaces, act, base="srnn", base_kwargs=cfg)
    policy.load_state_dict({k: torch.from_numpy(w[k]) for k in w.files})
    policy = policy.to(DEV)
    logger = logging.getLogger("crowdnav_eval_test")
    logger.setLevel(logging.INFO)
    cap = _Capture()
    logger.addHandler(cap)
    envs = make_vec_envs(cfg.env.env_name, cfg.env.seed, 1, cfg.reward.gamma, None, DEV, False, config=cfg)
    raw, disc, d2g = evaluate(actor_critic=policy, ob_rms=False, eval_envs=envs, num_processes=1, device=DEV, config=cfg,
                              logging=logger, visualize=False, recurrent_type="GRU")
    ref = evaluate_batched(policy, cfg, DEV)
    text = "\n".join(cap.lines)
    assert cap.lines[0] == "TEST"
    assert "success rate: %.3f" % ref["success"] in text and "collision rate: %.3f" % ref["collision"] in text
    assert "timeout rate: %.3f" % ref["timeout"] in text
    for must in ("Success cases: ", "SCENARIO BREAKDOWN: ", "SUCCESS CASES: ", "COLLISION CASES: ", "TIMEOUT CASES: ", "total: ",
                 "navigation time ======", "path length ======", "discounted reward ======", "non-discounted rewards ======",
                 "cumulative heading change ======", "SM1 - personal space violation ======", "SM5 - speed violation ======",
                 "MEAN: ", "STD DEV: ", "CI: ["):
        assert must in text, must
    assert set(raw) == set(disc) == set(d2g) == {"success", "collision", "timeout"}
    assert sum(len(v) for v in raw.values()) == 256 and len(raw["success"]) == round(ref["success"] * 256)
    mean_ret = np.mean([x[0] for v in raw.values() for x in v])
    assert abs(mean_ret - ref["mean_return"]) < 1e-4
    assert all(abs(a[0]) <= abs(b[0]) + 1e-6 for a, b in zip(disc["success"], raw["success"]))     # discounting shrinks the goal reward

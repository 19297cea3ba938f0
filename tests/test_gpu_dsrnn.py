"""GPU parity of the DS-RNN forward (K3) through the C ABI / Policy.act against the golden outputs of the
reference's own Policy.act on both shipped checkpoints (tests/golden/dsrnn_*.npz; tolerance 1e-3 relative,
BASELINE.json north_star) and against the torch fp32 restatement at a larger batch."""
import os

import numpy as np
import pytest
import torch

from crowdnav_dsrnn_b200 import Config
from crowdnav_dsrnn_b200.model import Policy
from crowdnav_dsrnn_b200.spaces import crowd_spaces
from oracle import dsrnn_oracle
from helpers import DSRNN_CASES, GOLDEN, TOL_NET_REL

pytestmark = pytest.mark.gpu
PRECISIONS = ["fp32", "bf16x3"]


def _policy(H, weights):
    obs, act = crowd_spaces(H)
    p = Policy(obs.spaces, act, base="srnn", base_kwargs=Config(human_num=H))
    sd = {k: torch.from_numpy(v) for k, v in weights.items()}
    p.load_state_dict(sd)
    return p.cuda(), sd


def _rel_err(got, want):
    want = np.asarray(want, np.float64)
    got = np.asarray(got, np.float64).reshape(want.shape)
    return np.abs(got - want).max() / max(1.0, np.abs(want).max())


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("case", DSRNN_CASES)
def test_forward_matches_reference_policy(case, prec):
    ckpt, h = case.rsplit("_h", 1)
    H = int(h)
    d = np.load(os.path.join(GOLDEN, "dsrnn_%s.npz" % case))
    w = np.load(os.path.join(GOLDEN, "weights_%s.npz" % ckpt))
    policy, _ = _policy(H, {k: w[k] for k in w.files})
    policy.precision = prec
    t = lambda k: torch.from_numpy(d[k]).cuda()
    obs = {"robot_node": t("robot_node"), "temporal_edges": t("temporal_edges"), "spatial_edges": t("spatial_edges")}
    hx = {"human_node_rnn": t("h_node"), "human_human_edge_rnn": t("h_edge")}
    value, mean, feat, hn, he = policy.cuda_forward(obs, hx, t("masks"))
    assert _rel_err(value.cpu().numpy(), d["ref_value"]) <= TOL_NET_REL
    assert _rel_err(mean.cpu().numpy(), d["ref_action_mean"]) <= TOL_NET_REL
    assert _rel_err(feat.cpu().numpy(), d["ref_actor_features"]) <= TOL_NET_REL
    assert _rel_err(hn.cpu().numpy(), d["ref_h_node"]) <= TOL_NET_REL
    assert _rel_err(he.cpu().numpy(), d["ref_h_edge"]) <= TOL_NET_REL
    # Policy.act: deterministic action == mean, log-prob of the mean, dict mutated in place
    v2, a2, lp, hx2 = policy.act(obs, hx, t("masks"), deterministic=True)
    assert hx2 is hx and hx["human_node_rnn"].shape == (d["h_node"].shape[0], 1, 128)
    assert _rel_err(a2.cpu().numpy(), d["ref_action_mean"]) <= TOL_NET_REL
    assert np.abs(lp.cpu().numpy() - d["ref_log_prob"]).max() <= 1e-3


@pytest.mark.parametrize("prec", PRECISIONS)
# (20, 4000): 329 tile pairs for 74 CTA pairs -> every persistent CTA of the edge kernel walks several tiles (barrier phases wrap);
# (1, 20000): 157 node tiles for 148 CTAs; (20, 129): odd tile counts in both problems; (3, 1): a single env
@pytest.mark.parametrize("H,n", [(5, 1000), (20, 777), (1, 130), (32, 65), (20, 4000), (1, 20000), (20, 129), (3, 1)])
def test_forward_matches_torch_restatement(H, n, prec):
    sd = dsrnn_oracle.random_state_dict(seed=H)
    policy, sd = _policy(H, {k: v.numpy() for k, v in sd.items()})
    policy.precision = prec
    g = torch.Generator().manual_seed(n)
    rn = torch.randn(n, 1, 7, generator=g) * 3
    te = torch.randn(n, 1, 2, generator=g)
    se = torch.randn(n, H, 2, generator=g) * 4
    hn = torch.randn(n, 1, 128, generator=g) * 0.5
    he = torch.randn(n, H + 1, 256, generator=g) * 0.5
    mk = (torch.rand(n, 1, generator=g) > 0.3).float()
    ref = dsrnn_oracle.forward(sd, rn, te, se, hn, he, mk)
    obs = {"robot_node": rn.cuda(), "temporal_edges": te.cuda(), "spatial_edges": se.cuda()}
    value, mean, feat, hn2, he2 = policy.cuda_forward(obs, {"human_node_rnn": hn.cuda(), "human_human_edge_rnn": he.cuda()}, mk.cuda())
    assert _rel_err(value.cpu().numpy(), ref["value"].numpy()) <= TOL_NET_REL
    assert _rel_err(mean.cpu().numpy(), ref["action_mean"].numpy()) <= TOL_NET_REL
    assert _rel_err(he2.cpu().numpy(), ref["h_edge"].numpy()) <= TOL_NET_REL
    assert _rel_err(hn2.cpu().numpy(), ref["h_node"].numpy()) <= TOL_NET_REL


def test_stochastic_act_log_prob_is_the_diag_gaussian_density():
    """act(deterministic=False): log-prob returned with the sample equals Normal(mean, exp(logstd)).log_prob(action).sum(-1)
    (distributions.py:36-38), and the samples have the right spread."""
    sd = dsrnn_oracle.random_state_dict(seed=3)
    policy, _ = _policy(5, {k: v.numpy() for k, v in sd.items()})
    with torch.no_grad():
        policy.dist.logstd._bias.copy_(torch.tensor([[-0.3], [0.2]], device="cuda"))
    n = 4096
    g = torch.Generator().manual_seed(11)
    obs = {"robot_node": torch.randn(n, 1, 7, generator=g).cuda(), "temporal_edges": torch.randn(n, 1, 2, generator=g).cuda(),
           "spatial_edges": torch.randn(n, 5, 2, generator=g).cuda()}
    hx = lambda: {"human_node_rnn": torch.zeros(n, 1, 128).cuda(), "human_human_edge_rnn": torch.zeros(n, 6, 256).cuda()}
    masks = torch.ones(n, 1).cuda()
    _, mean, _, _ = policy.act(obs, hx(), masks, deterministic=True)
    torch.manual_seed(5)
    _, action, lp, _ = policy.act(obs, hx(), masks, deterministic=False)
    std = policy.dist.logstd._bias.detach().view(1, 2).exp()
    want = torch.distributions.Normal(mean, std.expand_as(mean)).log_prob(action).sum(-1, keepdim=True)
    assert lp.shape == (n, 1) and (lp - want).abs().max() < 1e-4
    z = (action - mean) / std
    assert abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1.0) < 0.05


def test_cpu_tensors_fail_loudly():
    policy, _ = _policy(5, {k: v.numpy() for k, v in dsrnn_oracle.random_state_dict(1).items()})
    obs = {"robot_node": torch.zeros(2, 1, 7), "temporal_edges": torch.zeros(2, 1, 2), "spatial_edges": torch.zeros(2, 5, 2)}
    with pytest.raises(Exception):
        policy.act(obs, {"human_node_rnn": torch.zeros(2, 1, 128), "human_human_edge_rnn": torch.zeros(2, 6, 256)}, torch.ones(2, 1))


@pytest.mark.parametrize("prec,tol", [("fp16", 4e-3), ("bf16", 3e-2)])
def test_single_pass_modes_are_sane_but_not_the_parity_mode(prec, tol):
    """fp16 / bf16 run ONE tensor-core pass; measured worst relative errors (tools/precision_errors.py) are
    1.6e-3 / 1.1e-2, i.e. they miss the 1e-3 bar -- which is why bf16x3 is the default and the benchmarked mode."""
    d = np.load(os.path.join(GOLDEN, "dsrnn_holonomic_27776_h20.npz"))
    w = np.load(os.path.join(GOLDEN, "weights_holonomic_27776.npz"))
    policy, _ = _policy(20, {k: w[k] for k in w.files})
    policy.precision = prec
    t = lambda k: torch.from_numpy(d[k]).cuda()
    obs = {"robot_node": t("robot_node"), "temporal_edges": t("temporal_edges"), "spatial_edges": t("spatial_edges")}
    value, mean, feat, hn, he = policy.cuda_forward(obs, {"human_node_rnn": t("h_node"), "human_human_edge_rnn": t("h_edge")}, t("masks"))
    assert _rel_err(value.cpu().numpy(), d["ref_value"]) <= tol
    assert _rel_err(mean.cpu().numpy(), d["ref_action_mean"]) <= tol
    assert _rel_err(he.cpu().numpy(), d["ref_h_edge"]) <= tol

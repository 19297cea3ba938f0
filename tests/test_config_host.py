"""CPU: host-side logic -- Config restatement, step tables, CnConfig flattening, lazy infos, Policy container."""
import math

import numpy as np
import pytest
import torch

from crowdnav_dsrnn_b200 import Config, _lib, abi
from crowdnav_dsrnn_b200.envs import LazyInfos
from crowdnav_dsrnn_b200.info import Collision, Danger, Nothing, ReachGoal, Timeout
from crowdnav_dsrnn_b200.model import Policy
from crowdnav_dsrnn_b200.spaces import crowd_spaces
from helpers import GOLDEN, config_from_overrides
from oracle import dsrnn_oracle, ref_import


def test_config_defaults_and_derived_values():
    c = Config()
    assert (c.sim.human_num, c.sim.circle_radius, c.env.time_step, c.env.time_limit, c.env.test_size) == (5, 6, 0.25, 50, 500)
    assert c.reward.discomfort_penalty_factor == 10 * 0.25 and c.reward.collision_penalty == -20
    assert c.humans.random_goal_changing and c.humans.end_goal_changing and not c.robot.visible
    s = Config(social_metrics=True)
    assert (s.sim.circle_radius, s.env.test_size, s.test.side_preference) == (4, 2000, False)
    p = Config(test_sim=["side_pref_passing"])
    assert p.test.side_preference and p.sim.human_num == 1 and p.env.test_size == 200 and not p.humans.end_goal_changing


def test_step_tables_replay_float64_accumulation():
    t, words = abi.step_tables(0.25, 50)
    assert t == 196                                           # Timeout is returned by the 197th step
    steps = [s for s in range(300) if (words[s >> 5] >> (s & 31)) & 1]
    assert steps == [20, 40, 60, 80, 100, 120, 140, 160, 180]
    t, words = abi.step_tables(0.1, 50)
    assert t == 490 and not any(words)                        # dt = 0.1: `global_time % 5 == 0` never fires


def test_flatten_config_rejects_out_of_scope_flags():
    for sec, attr, val in (("humans", "policy", "cadrl"), ("reward", "norm_zones", True), ("lidar", "enable", True)):
        c = Config()
        setattr(getattr(c, sec), attr, val)
        with pytest.raises(NotImplementedError):
            abi.flatten_config(c, 4)
    c = Config()
    c.noise.add_noise = True                                  # no effect on the dict observation (crowd_sim_dict.py:71-103)
    assert bytes(abi.flatten_config(c, 4)) == bytes(abi.flatten_config(Config(), 4))
    c = Config()
    c.sim.group_human = True                                  # supported since round 2
    assert abi.flatten_config(c, 4).group_human == 1
    c = Config(test_sim=["side_pref_passing"])
    c.sim.group_human = True                                  # ... and disabled while testing side preferences (crowd_sim.py:123-125)
    assert abi.flatten_config(c, 4, phase="test").group_human == 0
    c = Config()
    c.humans.policy = "social_force"
    assert abi.flatten_config(c, 4).human_policy == abi.POLICY_SOCIAL_FORCE
    c = Config()
    c.sim.train_val_sim = "circle_crossing"
    with pytest.raises(TypeError):                            # crowd_sim.py:138-142
        abi.flatten_config(c, 4)
    c = Config(human_num=33)
    with pytest.raises(ValueError):
        abi.flatten_config(c, 4)


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree only exists in the build container")
def test_flatten_config_equals_reference_config():
    import json
    import glob
    import os
    from oracle import ref_harness

    for path in glob.glob(os.path.join(GOLDEN, "step_*.npz")):
        over = json.loads(str(np.load(path)["overrides"]))
        mine = abi.flatten_config(config_from_overrides(over), 32, phase="train")
        ref = abi.flatten_config(ref_harness.make_reference_config(**over), 32, phase="train")
        assert bytes(mine) == bytes(ref), path
    assert bytes(abi.flatten_config(Config(), 1)) == bytes(abi.flatten_config(ref_harness.make_reference_config(), 1))


class _FakeBuf:
    """A StepBuffers stand-in on the CPU: the same info block layout, filled by hand."""

    def __init__(self):
        from crowdnav_dsrnn_b200.engine import StepBuffers
        n = self.n = 3
        self.info_block = torch.zeros(StepBuffers.info_block_bytes(n), dtype=torch.uint8)
        for name, view in StepBuffers.carve_info(self.info_block, n).items():
            setattr(self, name, view)
        self.event.copy_(torch.tensor([abi.EV_NOTHING, abi.EV_DANGER, abi.EV_COLLISION], dtype=torch.int32))
        self.scenario.copy_(torch.tensor([0, 2, 3], dtype=torch.int32))
        self.info[1, abi.INFO_COLUMNS["dmin"]] = 0.125
        self.info[:, abi.INFO_COLUMNS["aggregate_nav_time"]] = torch.tensor([6.0, 5.0, 4.0])
        self.done.copy_(torch.tensor([0, 0, 1], dtype=torch.uint8))
        self.episode_return.copy_(torch.tensor([0.0, 0.0, -17.5]))
        self.episode_length.copy_(torch.tensor([0, 0, 42], dtype=torch.int32))


def test_lazy_infos_reference_shape():
    infos = LazyInfos(_FakeBuf(), side_preference=False, t0=0.0)
    assert len(infos) == 3
    a, b, c = list(infos)
    assert isinstance(a["info"]["event"], Nothing) and "episode" not in a
    assert isinstance(b["info"]["event"], Danger) and b["info"]["event"].min_dist == pytest.approx(0.125)
    assert b["info"]["scenario"] == "parallel_traffic" and b["info"]["aggregate_nav_time"] == 5
    assert isinstance(c["info"]["event"], Collision) and c["episode"]["r"] == -17.5 and c["episode"]["l"] == 42
    assert str(Timeout()) == "Timeout" and str(ReachGoal()) == "Reaching goal"
    assert "bad_transition" not in c


def test_info_block_layout_and_host_range():
    """reward | done sit next to each other at the end of the info block (one host copy in CrowdVecEnv.step_wait); every field
    starts on a 256-byte boundary; LazyInfos snapshots the block and carves its views on first use."""
    from crowdnav_dsrnn_b200.engine import StepBuffers
    for n in (1, 3, 16, 1000):
        fields, total = StepBuffers.info_layout(n)
        assert [f[0] for f in fields][-2:] == ["reward", "done"] and total == StepBuffers.info_block_bytes(n)
        assert all(f[3] % 256 == 0 for f in fields) and total % 256 == 0
        lo, hi, done_off = StepBuffers.host_range(n)
        by = {f[0]: f for f in fields}
        assert lo == by["reward"][3] and lo + done_off == by["done"][3] and hi == by["done"][3] + n and hi <= total
        block = (torch.arange(total) % 251).to(torch.uint8)
        views = StepBuffers.carve_info(block, n)
        host = block[lo:hi]
        assert torch.equal(host[:4 * n].view(torch.float32), views["reward"]) and torch.equal(host[done_off:done_off + n], views["done"])
        assert views["info"].shape == (n, abi.INFO_DIM) and views["event"].dtype is torch.int32
    buf = _FakeBuf()
    infos = LazyInfos(buf, side_preference=False, t0=0.0)
    assert infos._tensors is None and len(infos) == 3          # nothing carved yet
    buf.event.zero_()                                           # the engine reuses its buffer two steps later ...
    assert infos.tensors["event"].tolist() == [abi.EV_NOTHING, abi.EV_DANGER, abi.EV_COLLISION]      # ... the snapshot does not care
    assert infos.tensors is infos.tensors


def test_policy_state_dict_is_checkpoint_compatible_and_cpu_act_fails_loudly():
    obs, act = crowd_spaces(5)
    p = Policy(obs.spaces, act, base="srnn", base_kwargs=Config())
    w = np.load(GOLDEN + "/weights_holonomic_27776.npz")
    assert sorted(p.state_dict().keys()) == sorted(w.files) and len(w.files) == 45
    p.load_state_dict({k: torch.from_numpy(w[k]) for k in w.files})
    assert sum(v.numel() for v in p.parameters()) == 973983
    inputs = {"robot_node": torch.zeros(2, 1, 7), "temporal_edges": torch.zeros(2, 1, 2), "spatial_edges": torch.zeros(2, 5, 2)}
    hx = {"human_node_rnn": torch.zeros(2, 1, 128), "human_human_edge_rnn": torch.zeros(2, 6, 256)}
    with pytest.raises(_lib.CrowdNavLibraryError):
        p.act(inputs, hx, torch.ones(2, 1))


def test_evaluate_actions_matches_reference_forward_on_cpu():
    """The differentiable training path restates the same math: T=1 chunk vs the golden reference outputs."""
    d = np.load(GOLDEN + "/dsrnn_holonomic_27776_h5.npz")
    w = np.load(GOLDEN + "/weights_holonomic_27776.npz")
    obs, act = crowd_spaces(5)
    p = Policy(obs.spaces, act, base="srnn", base_kwargs=Config())
    p.load_state_dict({k: torch.from_numpy(w[k]) for k in w.files})
    t = lambda k: torch.from_numpy(d[k])
    inputs = {"robot_node": t("robot_node"), "temporal_edges": t("temporal_edges"), "spatial_edges": t("spatial_edges")}
    hx = {"human_node_rnn": t("h_node"), "human_human_edge_rnn": t("h_edge")}
    value, logp, ent, hx2 = p.evaluate_actions(inputs, hx, t("masks"), t("ref_action_mean"))
    assert (value - t("ref_value")).abs().max() < 1e-4
    assert (logp - t("ref_log_prob")).abs().max() < 1e-4
    assert (hx2["human_human_edge_rnn"] - t("ref_h_edge")).abs().max() < 1e-5
    assert value.requires_grad and math.isfinite(float(ent))


def _sequence_case(dtype, device, n=6, H=5, T=7, seed=0):
    g = torch.Generator().manual_seed(seed)
    obs_space, act_space = crowd_spaces(H)
    # the parameters come from the GLOBAL generator: seed it here (and restore it), otherwise the draw -- and with it how close
    # the comparison comes to its tolerance -- depends on which tests ran before this one
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(1234 + seed)
        policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=Config(human_num=H))
    policy = policy.to(device=device, dtype=dtype)
    mk = lambda *shape, scale=1.0: (torch.randn(*shape, generator=g) * scale).to(device=device, dtype=dtype)
    obs = {"robot_node": mk(T * n, 1, 7), "temporal_edges": mk(T * n, 1, 2), "spatial_edges": mk(T * n, H, 2, scale=3.0)}
    masks = (torch.rand(T * n, 1, generator=g) > 0.2).to(device=device, dtype=dtype)
    action = mk(T * n, 2)
    h0 = {"human_node_rnn": mk(n, 1, 128, scale=0.3), "human_human_edge_rnn": mk(n, H + 1, 256, scale=0.3)}
    return policy, obs, masks, action, h0


def sequence_impls_agree(dtype, device, tol, impls=("per_step", "batched"), case=None):
    """evaluate_actions over a [T, n] chunk: the batched form (one masked GRU sequence per recurrent unit, everything else
    evaluated for all T*n samples at once) against the per-step autograd graph -- outputs, final hidden state, every
    parameter gradient and the gradient with respect to the initial hidden state."""
    policy, obs, masks, action, h0 = _sequence_case(dtype, device, **(case or {}))
    res = {}
    for impl in impls:
        policy.sequence_impl = impl
        policy.zero_grad(set_to_none=True)
        hx = {k: v.clone().requires_grad_(True) for k, v in h0.items()}
        value, logp, _, out = policy.evaluate_actions(obs, dict(hx), masks, action)
        weights = torch.linspace(0, 1, value.numel(), device=value.device, dtype=dtype).view_as(value)
        ((value * weights).sum() + 0.3 * logp.sum() + 0.01 * out["human_human_edge_rnn"].sum()
         + 0.02 * out["human_node_rnn"].sum()).backward()
        grads = {k: q.grad.clone() for k, q in policy.named_parameters() if q.grad is not None}
        res[impl] = (value.detach(), logp.detach(), out["human_human_edge_rnn"].detach(), out["human_node_rnn"].detach(),
                     grads, [hx[k].grad.clone() for k in sorted(hx)])
    a, b = res[impls[0]], res[impls[1]]
    for x, y in list(zip(a[:4], b[:4])) + list(zip(a[5], b[5])):
        assert (x - y).abs().max().item() <= tol * max(1.0, x.abs().max().item())
    assert sorted(a[4]) == sorted(b[4]) and len(a[4]) >= 40
    bad = []
    for k in a[4]:
        scale = max(a[4][k].abs().max().item(), 1e-3)
        err = (a[4][k] - b[4][k]).abs().max().item()
        if not err <= tol * scale:
            bad.append("%s: err %.3e scale %.3e" % (k, err, scale))
    assert not bad, bad


def test_batched_sequence_forward_equals_per_step_graph_on_cpu():
    sequence_impls_agree(torch.float64, "cpu", 1e-10)
    sequence_impls_agree(torch.float32, "cpu", 2e-4)


def test_policy_can_be_pickled_and_deep_copied():
    """train.py's resume path loads a pickled actor_critic (train.py:170-172): library handles and caches must stay behind."""
    import copy
    import io

    obs, act = crowd_spaces(5)
    p = Policy(obs.spaces, act, base="srnn", base_kwargs=Config())
    p.__dict__["_param_list"] = p._weight_tensors()          # as after a forward
    q = copy.deepcopy(p)
    buf = io.BytesIO()
    torch.save(p, buf)
    buf.seek(0)
    r = torch.load(buf, weights_only=False)
    for other in (q, r):
        assert other._handle is None and "_param_list" not in other.__dict__
        assert all(torch.equal(a, b) for a, b in zip(p.state_dict().values(), other.state_dict().values()))

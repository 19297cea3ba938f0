"""GPU: the CUDA-graph rollout (rollout.py) replays exactly what the eager Policy.act -> venv.step loop computes."""
import numpy as np
import pytest
import torch

from crowdnav_dsrnn_b200 import Config
from crowdnav_dsrnn_b200.envs import CrowdVecEnv
from crowdnav_dsrnn_b200.model import Policy
from crowdnav_dsrnn_b200.rollout import GraphedRollout
from crowdnav_dsrnn_b200.spaces import crowd_spaces
from oracle import dsrnn_oracle

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("edge_image", [True, False])
@pytest.mark.parametrize("H,n,kin", [(5, 64, "holonomic"), (10, 300, "unicycle")])
def test_graph_replay_equals_eager_loop(H, n, kin, edge_image):
    """edge_image=True (the default): every forward of the replayed loop reads its hidden-state operand by TMA from the
    split-bf16 image the previous forward's epilogue wrote (cn_dsrnn_set_edge_image) -- same bits as converting the fp32 state."""
    cfg = Config(kinematics=kin, human_num=H)
    obs_space, act_space = crowd_spaces(H)
    policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg)
    policy.load_state_dict(dsrnn_oracle.random_state_dict(3), strict=True)
    policy = policy.to(DEV)
    steps = 40

    def eager():
        venv = CrowdVecEnv(cfg, n, DEV, seed=5, phase="train")
        obs = venv.reset()
        hx = {"human_node_rnn": torch.zeros(n, 1, 128, device=DEV), "human_human_edge_rnn": torch.zeros(n, H + 1, 256, device=DEV)}
        masks = torch.zeros(n, 1, device=DEV)
        trace = []
        for _ in range(steps):
            _, action, _, hx = policy.act(obs, hx, masks, deterministic=True)
            obs, reward, done, buf = venv.step_device(action)
            masks = (1.0 - done.float()).unsqueeze(1)
            trace.append((reward.clone(), done.clone(), buf.event.clone()))
        st = {k: v.clone() for k, v in venv.engine.get_state().items()}
        venv.close()
        return trace, st, hx["human_human_edge_rnn"].clone()

    def graphed():
        venv = CrowdVecEnv(cfg, n, DEV, seed=5, phase="train")
        obs = venv.reset()
        roll = GraphedRollout(policy, venv, obs, edge_image=edge_image)      # the constructor already plays 2 (eager) steps
        assert roll.use_image == edge_image
        for _ in range(steps - roll.steps):
            buf = roll.step()
        st = {k: v.clone() for k, v in venv.engine.get_state().items()}
        hx, _ = roll.hidden()
        out = (buf.reward.clone(), buf.done.clone(), buf.event.clone()), st, hx["human_human_edge_rnn"].clone()
        venv.close()
        return out

    trace, st_e, he_e = eager()
    last_g, st_g, he_g = graphed()
    for a, b in zip(trace[-1], last_g):
        assert torch.equal(a, b)
    for k in st_e:
        assert torch.equal(st_e[k], st_g[k]), k
    assert torch.equal(he_e, he_g)
    assert int(sum(t[1].sum() for t in trace)) > 0          # episodes ended and were re-spawned inside the graphs too


@pytest.mark.parametrize("H,n,split,kin", [(5, 96, 40, "holonomic"), (10, 300, None, "unicycle")])
def test_pipelined_half_batches_equal_eager_loop(H, n, split, kin):
    """rollout.PipelinedRollout: two half batches on two streams in one graph leave every env exactly where the plain loop
    does (env state bit for bit; half 0's hidden state after c forwards, half 1's after c + 1)."""
    from crowdnav_dsrnn_b200.rollout import PipelinedRollout

    cfg = Config(kinematics=kin, human_num=H)
    obs_space, act_space = crowd_spaces(H)
    policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg)
    policy.load_state_dict(dsrnn_oracle.random_state_dict(3), strict=True)
    policy = policy.to(DEV)
    cycles = 40

    venv = CrowdVecEnv(cfg, n, DEV, seed=5, phase="train")
    obs = venv.reset()
    hx = {"human_node_rnn": torch.zeros(n, 1, 128, device=DEV), "human_human_edge_rnn": torch.zeros(n, H + 1, 256, device=DEV)}
    masks = torch.zeros(n, 1, device=DEV)
    dones = 0
    for _ in range(cycles):
        _, action, _, hx = policy.act(obs, hx, masks, deterministic=True)
        obs, reward, done, buf = venv.step_device(action)
        masks = (1.0 - done.float()).unsqueeze(1)
        dones += int(done.sum())
    st_e = {k: v.clone() for k, v in venv.engine.get_state().items()}
    he_c = hx["human_human_edge_rnn"].clone()
    hn_c = hx["human_node_rnn"].clone()
    value, action, _, hx = policy.act(obs, hx, masks, deterministic=True)
    he_c1, hn_c1 = hx["human_human_edge_rnn"].clone(), hx["human_node_rnn"].clone()
    venv.close()
    assert dones > 0

    roll = PipelinedRollout(policy, cfg, n, DEV, seed=5, phase="train", split=split)
    for _ in range(cycles - roll.steps):
        roll.step()
    torch.cuda.synchronize()
    halves = roll.state()
    n_a = roll.sizes[0]
    for k in st_e:
        got = torch.cat([h["engine"].get_state()[k] for h in halves], 0)
        assert torch.equal(st_e[k], got), k
    assert torch.equal(halves[0]["hidden"]["human_human_edge_rnn"], he_c[:n_a])
    assert torch.equal(halves[0]["hidden"]["human_node_rnn"], hn_c[:n_a])
    assert torch.equal(halves[1]["hidden"]["human_human_edge_rnn"], he_c1[n_a:])
    assert torch.equal(halves[1]["hidden"]["human_node_rnn"], hn_c1[n_a:])
    assert torch.equal(halves[1]["action"], action[n_a:]) and torch.equal(halves[1]["value"], value[n_a:])
    roll.close()

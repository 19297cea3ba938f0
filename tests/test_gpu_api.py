"""GPU: the reference-facing Python API (test plan T6): make_vec_envs / VecEnv contract, lazy infos,
auto-reset bookkeeping, the single-env CrowdSimDict view, and a train.py-style rollout loop (train.py:226-292)."""
import numpy as np
import pytest
import torch

from crowdnav_dsrnn_b200 import Config
from crowdnav_dsrnn_b200.crowd_sim_dict import CrowdSimDict
from crowdnav_dsrnn_b200.envs import make_vec_envs
from crowdnav_dsrnn_b200.info import Collision, Danger, Nothing, ReachGoal, Timeout
from crowdnav_dsrnn_b200.model import Policy

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def test_make_vec_envs_contract_and_train_style_rollout():
    cfg = Config()
    n = 16
    cfg.training.num_processes = n
    envs = make_vec_envs(cfg.env.env_name, cfg.env.seed, n, cfg.reward.gamma, None, DEV, False, config=cfg)
    assert envs.num_envs == n and envs.action_space.shape == (2,)
    spaces = envs.observation_space.spaces
    assert spaces["robot_node"].shape == (1, 7) and spaces["temporal_edges"].shape == (1, 2) and spaces["spatial_edges"].shape == (5, 2)
    policy = Policy(spaces, envs.action_space, base=cfg.robot.policy, base_kwargs=cfg).to(DEV)
    assert policy.is_recurrent and policy.base.human_num == 5 and policy.base.nenv == n
    obs = envs.reset()
    for k, shape in (("robot_node", (n, 1, 7)), ("temporal_edges", (n, 1, 2)), ("spatial_edges", (n, 5, 2))):
        assert obs[k].shape == shape and obs[k].dtype == torch.float32 and obs[k].device == DEV
    assert torch.equal(obs["temporal_edges"], torch.zeros(n, 1, 2, device=DEV))          # robot starts at rest
    hx = {"human_node_rnn": torch.zeros(n, 1, 128, device=DEV), "human_human_edge_rnn": torch.zeros(n, 6, 256, device=DEV)}
    masks = torch.zeros(n, 1, device=DEV)
    events, episodes = set(), 0
    for step in range(260):
        with torch.no_grad():
            value, action, logp, hx = policy.act(obs, hx, masks)
        assert value.shape == (n, 1) and action.shape == (n, 2) and logp.shape == (n, 1)
        prev = obs
        obs, reward, done, infos = envs.step(action)
        assert reward.shape == (n, 1) and reward.dtype == torch.float32 and reward.device.type == "cpu"
        assert isinstance(done, np.ndarray) and done.dtype == bool and done.shape == (n,) and len(infos) == n
        for i, info in enumerate(infos):
            ev = info["info"]["event"]
            assert isinstance(ev, (Collision, Danger, Nothing, ReachGoal, Timeout))
            assert info["info"]["scenario"] in cfg.sim.train_val_sim and "bad_transition" not in info
            events.add(type(ev).__name__)
            assert ("episode" in info) == bool(done[i])
            if done[i]:
                episodes += 1
                assert set(info["episode"]) == {"r", "l", "t"} and info["episode"]["l"] >= 1
                assert isinstance(ev, (Collision, ReachGoal, Timeout))
                # auto-reset: the observation returned with a terminal step is the NEW episode's first observation
                assert torch.equal(obs["temporal_edges"][i], torch.zeros(1, 2, device=DEV))
        masks = torch.FloatTensor([[0.0] if d else [1.0] for d in done]).to(DEV)
        assert prev["robot_node"] is not obs["robot_node"]            # double-buffered: the previous observation survives
    assert episodes > 0 and {"Nothing"} <= events
    base_env = envs.venv.envs[0].env                                   # evaluation.py:71
    assert base_env.time_step == cfg.env.time_step and base_env.time_limit == cfg.env.time_limit
    assert base_env.robot.v_pref == cfg.robot.v_pref and base_env.global_time >= 0
    envs.close()


def test_single_env_crowd_sim_dict_view():
    cfg = Config(kinematics="unicycle", human_num=3)
    env = CrowdSimDict()
    with pytest.raises(AttributeError):
        env.reset()                                                    # "robot has to be set!" (crowd_sim_dict.py:132-133)
    env.configure(cfg)
    env.thisSeed, env.nenv, env.phase = 7, 1, "test"
    ob = env.reset()
    assert ob["robot_node"].shape == (1, 7) and ob["spatial_edges"].shape == (3, 2)
    total, done, steps = 0.0, False, 0
    while not done and steps < 600:
        ob, reward, done, info = env.step(np.array([0.05, 0.01], dtype=np.float32))
        assert isinstance(reward, float) and isinstance(done, bool) and "episode" not in info
        total += reward
        steps += 1
    assert done and env.global_time == pytest.approx(steps * cfg.env.time_step)
    assert type(info["info"]["event"]).__name__ in ("Collision", "ReachGoal", "Timeout")
    env.close()


def test_string_scenario_list_is_rejected_like_the_reference():
    cfg = Config()
    cfg.sim.train_val_sim = "circle_crossing"
    with pytest.raises(TypeError):
        make_vec_envs(cfg.env.env_name, 0, 4, 0.99, None, DEV, False, config=cfg)

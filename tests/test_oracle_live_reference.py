"""CPU, build container only: the C oracle against the REFERENCE ITSELF (its own CrowdSimDict.step imported from
/root/reference under the oracle/ref_import.py shims) on FRESH seeds -- the committed golden vectors are one draw of this
comparison (oracle/gen_golden.py); skipped where the reference is not mounted (the GPU box)."""
import pytest

from oracle import crowd_oracle, gen_golden, ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference is not mounted here")


@pytest.mark.parametrize("name,seed", [("c1_holonomic_h5", 9001), ("c2_unicycle_h10", 9002), ("c3_fov_h20", 9003),
                                       ("c4_social_h5", 9004), ("x_robot_visible_h5", 9005), ("x_human_fov_h6", 9006)])
def test_oracle_matches_reference_on_fresh_seeds(name, seed):
    crowd_oracle.build()
    over, _, _ = gen_golden.CASES[name]
    _, _, ref, rep = gen_golden.reference_vs_oracle(over, 64, seed)
    assert rep["flag_mismatch"] == 0 and rep["vis_mismatch"] == 0 and rep["int_info_mismatch"] == 0, rep
    assert rep["reward"] <= 1e-5 and rep["robot_pv"] <= 1e-4 and rep["human_pv"] <= 1e-4, rep
    assert rep["belief"] <= 1e-4 and rep["obs"] <= 1e-4 and rep["extras"] <= 1e-4 and rep["jerk"] <= 1e-4, rep
    assert ref["done"].shape == (64,)


@pytest.mark.parametrize("ckpt", ["holonomic_27776", "unicycle_55554"])
@pytest.mark.parametrize("H,seed", [(3, 77), (20, 78)])
def test_dsrnn_restatement_matches_reference_policy_on_fresh_inputs(ckpt, H, seed):
    """oracle/dsrnn_oracle.py against the reference's own Policy.act (shipped checkpoints) on inputs that are not in the fixtures."""
    import os

    import numpy as np
    import torch

    from oracle import dsrnn_oracle, gen_golden_dsrnn, ref_harness

    ref_import.install_shims()
    from pytorchBaselines.a2c_ppo_acktr.model import Policy

    torch.set_num_threads(2)
    n = 16
    sd = torch.load(os.path.join(ref_import.REFERENCE_ROOT, gen_golden_dsrnn.CKPTS[ckpt]), map_location="cpu")
    cfg = ref_harness.make_reference_config(**{"sim.human_num": H, "training.cuda": False, "training.num_processes": n})
    spaces = {"robot_node": ref_import.Box(-np.inf, np.inf, (1, 7)), "temporal_edges": ref_import.Box(-np.inf, np.inf, (1, 2)),
              "spatial_edges": ref_import.Box(-np.inf, np.inf, (H, 2))}
    policy = Policy(spaces, ref_import.Box(-np.inf, np.inf, (2,)), base="srnn", base_kwargs=cfg)
    policy.load_state_dict(sd)
    policy.eval()
    rn, te, se, hn, he, mk = gen_golden_dsrnn.sample_inputs(n, H, seed)
    with torch.no_grad():
        value, action, _, hx = policy.act({"robot_node": rn, "temporal_edges": te, "spatial_edges": se},
                                          {"human_node_rnn": hn.clone(), "human_human_edge_rnn": he.clone()}, mk, deterministic=True)
    mine = dsrnn_oracle.forward(sd, rn, te, se, hn, he, mk)
    assert float((mine["value"] - value.reshape(mine["value"].shape)).abs().max()) <= 1e-4
    assert float((mine["action_mean"] - action.reshape(mine["action_mean"].shape)).abs().max()) <= 1e-4
    assert float((mine["h_edge"] - hx["human_human_edge_rnn"].reshape(mine["h_edge"].shape)).abs().max()) <= 1e-5
    assert float((mine["h_node"] - hx["human_node_rnn"].reshape(mine["h_node"].shape)).abs().max()) <= 1e-5


def test_ppo_update_matches_reference_on_a_fresh_rollout():
    """The PPO update path against the reference's own SRNNRolloutStorage + PPO.update on a rollout / permutation that is not
    the committed fixture (tests/test_ppo_update.py holds the helpers)."""
    import torch

    import test_ppo_update as T
    from crowdnav_dsrnn_b200.ppo import PPO
    from oracle import gen_golden_ppo

    g = gen_golden_ppo.reference_run(rollout_seed=1234, act_seed=21, perm_seed=777)
    hyper = g["hyper"]
    torch.set_num_threads(4)
    policy = T._policy()
    before = {k: v.detach().clone() for k, v in policy.state_dict().items()}
    st = T._storage(g)
    st.compute_returns(torch.from_numpy(g["next_value"]), True, float(hyper[8]), float(hyper[9]), False)
    agent = PPO(policy, float(hyper[0]), int(hyper[1]), int(hyper[2]), float(hyper[3]), float(hyper[4]), lr=float(hyper[5]),
                eps=float(hyper[6]), max_grad_norm=float(hyper[7]))
    torch.manual_seed(int(hyper[10]))
    T._check_update(g, policy, before, agent.update(st))

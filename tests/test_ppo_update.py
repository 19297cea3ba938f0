"""PPO update path against the reference's own SRNNRolloutStorage + PPO.update run on a fixed synthetic
rollout (tests/golden/ppo_update_h5.npz, made by oracle/gen_golden_ppo.py).  CPU-only: the update is torch-level code; the
2-rank data-parallel variant runs over gloo."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crowdnav_dsrnn_b200.config import Config
from crowdnav_dsrnn_b200.model import Policy
from crowdnav_dsrnn_b200.ppo import PPO, normalized_advantages, update_linear_schedule
from crowdnav_dsrnn_b200.spaces import crowd_spaces
from crowdnav_dsrnn_b200.storage import SRNNRolloutStorage

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
T, N, H = 30, 6, 5


def _golden():
    return np.load(os.path.join(GOLDEN, "ppo_update_h5.npz"))


def _policy():
    w = np.load(os.path.join(GOLDEN, "weights_holonomic_27776.npz"))
    obs, act = crowd_spaces(H)
    p = Policy(obs.spaces, act, base="srnn", base_kwargs=Config())
    p.load_state_dict({k: torch.from_numpy(w[k]) for k in w.files})
    return p


def _storage(g, envs=None, keep_hidden_history=True):
    """Rebuild the golden rollout through the public `insert` interface (optionally only the env columns `envs`)."""
    obs, act = crowd_spaces(H)
    envs = list(range(N)) if envs is None else list(envs)
    t = lambda k: torch.from_numpy(g[k])[:, envs] if g[k].shape[0] in (T, T + 1) else torch.from_numpy(g[k])[envs]
    st = SRNNRolloutStorage(T, len(envs), obs.spaces, act, 128, 256, "GRU", keep_hidden_history=keep_hidden_history)
    for k in st.obs:
        st.obs[k][0].copy_(t("obs_" + k)[0])
    st.masks[0].copy_(t("masks")[0])
    st.bad_masks[0].copy_(t("bad_masks")[0])
    st.recurrent_hidden_states["human_node_rnn"][0].copy_(t("h_node0"))
    st.recurrent_hidden_states["human_human_edge_rnn"][0].copy_(t("h_edge0"))
    junk = {"human_node_rnn": torch.zeros(len(envs), 1, 128), "human_human_edge_rnn": torch.zeros(len(envs), H + 1, 256)}
    for s in range(T):
        st.insert({k: t("obs_" + k)[s + 1] for k in st.obs}, junk, t("actions")[s], t("action_log_probs")[s],
                  t("value_preds")[s], t("rewards")[s], t("masks")[s + 1], t("bad_masks")[s + 1])
    assert st.step == 0
    return st


def _check_update(g, policy, before, losses, rtol=0.02, atol=2e-7):
    ref_losses = g["losses"]
    assert abs(losses[0] - ref_losses[0]) <= 1e-4 * abs(ref_losses[0])
    assert abs(losses[1] - ref_losses[1]) <= 1e-4
    assert abs(losses[2] - ref_losses[2]) <= 1e-5
    names = [str(x) for x in g["param_names"]]
    sd = policy.state_dict()
    assert sorted(sd) == names
    worst = 0.0
    for i, k in enumerate(names):
        delta = (sd[k].detach() - before[k]).double().reshape(-1)
        head = g["delta_head_%02d" % i]
        norm, total, amax = g["delta_stats_%02d" % i]
        err = np.abs(delta[:head.size].numpy() - head)
        assert (err <= rtol * np.abs(head) + atol).all(), (k, err.max())
        assert abs(float(delta.norm()) - norm) <= rtol * norm + atol, k
        assert abs(float(delta.abs().max()) - amax) <= rtol * amax + atol, k
        worst = max(worst, float(err.max()))
    return worst


@pytest.mark.parametrize("tag,use_gae,proper", [("gae", True, False), ("gae_proper", True, True), ("nogae", False, False),
                                                  ("nogae_proper", False, True)])
def test_compute_returns_matches_reference(tag, use_gae, proper):
    g = _golden()
    st = _storage(g)
    hyper = g["hyper"]
    st.compute_returns(torch.from_numpy(g["next_value"]), use_gae, float(hyper[8]), float(hyper[9]), proper)
    np.testing.assert_allclose(st.returns.numpy(), g["returns_" + tag], rtol=1e-6, atol=1e-6)


def test_advantages_and_generator_layout():
    g = _golden()
    st = _storage(g)
    st.returns.copy_(torch.from_numpy(g["returns_gae"]))
    adv = normalized_advantages(st)
    np.testing.assert_allclose(adv.numpy(), g["advantages"], rtol=1e-5, atol=1e-6)
    perm = torch.tensor([4, 1, 5, 0, 2, 3])
    batches = list(st.recurrent_generator(adv, 2, perm=perm))
    assert len(batches) == 2
    obs_b, hx_b, act_b, val_b, ret_b, mask_b, logp_b, adv_b = batches[1]
    assert obs_b["spatial_edges"].shape == (T * 3, H, 2) and hx_b["human_human_edge_rnn"].shape == (3, H + 1, 256)
    # time-major flattening of whole env trajectories (storage.py:264-290): row t*n+j is env perm[3+j] at step t
    for j, e in enumerate([0, 2, 3]):
        np.testing.assert_array_equal(act_b.view(T, 3, 2)[:, j].numpy(), g["actions"][:, e])
        np.testing.assert_array_equal(adv_b.view(T, 3, 1)[:, j].numpy(), adv[:, e].numpy())
        np.testing.assert_array_equal(hx_b["human_node_rnn"][j].numpy(), g["h_node0"][e])
        np.testing.assert_array_equal(obs_b["robot_node"].view(T, 3, 1, 7)[:, j].numpy(), g["obs_robot_node"][:T, e])
    with pytest.raises(AssertionError):
        list(st.recurrent_generator(adv, 7))


@pytest.mark.parametrize("per_pass,keep,flags", [(None, True, {}), (2, False, {}), (2, False, {"tf32": True, "bf16x3": True})])
def test_update_matches_reference(per_pass, keep, flags):
    """Losses and every parameter's change after 5 epochs x 2 minibatches equal the reference's PPO.update.  (The GEMM
    arithmetic switches only act on CUDA tensors; here they must be inert and restored after the update.)"""
    g = _golden()
    hyper = g["hyper"]
    torch.set_num_threads(4)
    policy = _policy()
    before = {k: v.detach().clone() for k, v in policy.state_dict().items()}
    st = _storage(g, keep_hidden_history=keep)
    st.compute_returns(torch.from_numpy(g["next_value"]), True, float(hyper[8]), float(hyper[9]), False)
    agent = PPO(policy, float(hyper[0]), int(hyper[1]), int(hyper[2]), float(hyper[3]), float(hyper[4]), lr=float(hyper[5]),
                eps=float(hyper[6]), max_grad_norm=float(hyper[7]), max_envs_per_pass=per_pass, **flags)
    torch.manual_seed(int(hyper[10]))          # the reference draws torch.randperm(N) per epoch from the CPU generator
    tf32_before = torch.backends.cuda.matmul.allow_tf32
    losses = agent.update(st)
    _check_update(g, policy, before, losses)
    from crowdnav_dsrnn_b200 import model as model_mod
    assert model_mod.SEQUENCE_GEMM == "fp32" and torch.backends.cuda.matmul.allow_tf32 == tf32_before
    st.after_update()
    np.testing.assert_array_equal(st.obs["robot_node"][0].numpy(), g["obs_robot_node"][-1])
    np.testing.assert_array_equal(st.masks[0].numpy(), g["masks"][-1])


def test_linear_schedule():
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1.0)
    update_linear_schedule(opt, 25, 100, 4e-5)
    assert opt.param_groups[0]["lr"] == pytest.approx(3e-5)


# ----------------------------------------------------------------------------- 2-rank data parallel == single process
def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        g = _golden()
        hyper = g["hyper"]
        envs = [0, 1, 2] if rank == 0 else [3, 4, 5]
        policy = _policy()
        before = {k: v.detach().clone() for k, v in policy.state_dict().items()}
        st = _storage(g, envs=envs)
        st.compute_returns(torch.from_numpy(g["next_value"])[envs], True, float(hyper[8]), float(hyper[9]), False)
        # the golden run's permutations, split so that every global minibatch is the union of the ranks' local ones
        torch.manual_seed(int(hyper[10]))
        perms = [torch.randperm(N) for _ in range(int(hyper[1]))]
        usable = [p for p in perms if all(sorted(int(e >= 3) for e in p[s:s + 3].tolist()) in ([0, 0, 0], [1, 1, 1]) for s in (0, 3))]
        agent = PPO(policy, float(hyper[0]), int(hyper[1]), 1, float(hyper[3]), float(hyper[4]), lr=float(hyper[5]),
                    eps=float(hyper[6]), max_grad_norm=float(hyper[7]))
        agent.perm_fn = lambda n: torch.arange(n)
        adv = normalized_advantages(st)
        np.testing.assert_allclose(adv.numpy(), g["advantages"][:, envs], rtol=1e-5, atol=1e-6)   # global statistics
        losses = agent.update(st)
        digest = torch.cat([v.detach().reshape(-1)[:8] for v in policy.state_dict().values()])
        gathered = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        assert torch.equal(gathered[0], gathered[1]), "ranks diverged"
        ret[rank] = (losses, agent.allreduce_calls, {k: (v.detach() - before[k]) for k, v in policy.state_dict().items()}, len(usable))
    finally:
        dist.destroy_process_group()


def _single_process_reference_run():
    """Single-process update with num_mini_batch=1 on all 6 envs: what 2 ranks x 3 envs must reproduce."""
    g = _golden()
    hyper = g["hyper"]
    policy = _policy()
    before = {k: v.detach().clone() for k, v in policy.state_dict().items()}
    st = _storage(g)
    st.compute_returns(torch.from_numpy(g["next_value"]), True, float(hyper[8]), float(hyper[9]), False)
    agent = PPO(policy, float(hyper[0]), int(hyper[1]), 1, float(hyper[3]), float(hyper[4]), lr=float(hyper[5]),
                eps=float(hyper[6]), max_grad_norm=float(hyper[7]))
    agent.perm_fn = lambda n: torch.arange(n)
    losses = agent.update(st)
    return losses, {k: (v.detach() - before[k]) for k, v in policy.state_dict().items()}


def test_two_rank_gloo_update_equals_single_process():
    port = 29600 + os.getpid() % 300
    manager = mp.get_context("spawn").Manager()
    ret = manager.dict()
    mp.spawn(_dp_worker, args=(2, port, ret), nprocs=2, join=True)
    ref_losses, ref_delta = _single_process_reference_run()
    for rank in (0, 1):
        losses, calls, delta, _ = ret[rank]
        # gradient all-reduces: one flat bucket for the first minibatch (which also records which parameters receive
        # gradients), then two per minibatch -- the early bucket is started from a gradient hook while backward is still
        # running, the edge-GRU bucket follows after backward
        assert calls == 1 + 2 * 4
        assert abs(losses[0] - ref_losses[0]) <= 1e-4 * abs(ref_losses[0])
        assert abs(losses[1] - ref_losses[1]) <= 1e-4
        for k in ref_delta:
            a, b = delta[k].double(), ref_delta[k].double()
            assert float((a - b).norm()) <= 0.02 * float(b.norm()) + 1e-7, k

"""TEST INFRASTRUCTURE -- golden vector for the PPO update path: runs the REFERENCE's own
`SRNNRolloutStorage` + `PPO.update` + `Policy.evaluate_actions` (pytorchBaselines/a2c_ppo_acktr/{storage.py,algo/ppo.py,
model.py}) on a fixed synthetic rollout, starting from the shipped checkpoint, and records the rollout inputs, the returns,
the three losses and a digest of every updated parameter.

Build container only (imports /root/reference under oracle/ref_import.py shims):

    python -m oracle.gen_golden_ppo          # writes tests/golden/ppo_update_h5.npz

The env permutation of each epoch is `torch.randperm(N)` on the CPU generator after `torch.manual_seed(PERM_SEED)`; the
restatement draws the same numbers the same way, so both sides see identical minibatches.
"""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
T, N, H = 30, 6, 5
PERM_SEED = 4242
HEAD = 48          # leading elements of each updated parameter kept verbatim


def synthetic_rollout(policy, storage, seed):
    """Fill `storage` by acting with `policy` on random observations (no simulator needed for the update arithmetic)."""
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, generator=g)
    obs = lambda: {"robot_node": rnd(N, 1, 7) * 2.0, "temporal_edges": rnd(N, 1, 2) * 0.5, "spatial_edges": rnd(N, H, 2) * 3.0}
    first = obs()
    for k in storage.obs:
        storage.obs[k][0].copy_(first[k])
    storage.masks[0].copy_((torch.rand(N, 1, generator=g) > 0.3).float())
    storage.recurrent_hidden_states["human_node_rnn"][0].copy_(rnd(N, 1, 128) * 0.3)
    storage.recurrent_hidden_states["human_human_edge_rnn"][0].copy_(rnd(N, H + 1, 256) * 0.3)
    for step in range(T):
        with torch.no_grad():
            value, action, logp, hx = policy.act({k: storage.obs[k][step] for k in storage.obs},
                                                 {k: storage.recurrent_hidden_states[k][step].clone() for k in storage.recurrent_hidden_states},
                                                 storage.masks[step])
        # perturb the stored action so that ratio != 1 from the first epoch on is not needed; keep the sampled one
        reward = rnd(N, 1) * 0.5
        masks = (torch.rand(N, 1, generator=g) > 0.08).float()
        bad = (torch.rand(N, 1, generator=g) > 0.03).float()
        storage.insert(obs(), hx, action, logp, value, reward, masks, bad)
    with torch.no_grad():
        next_value = policy.get_value({k: storage.obs[k][-1] for k in storage.obs},
                                      {k: storage.recurrent_hidden_states[k][-1].clone() for k in storage.recurrent_hidden_states},
                                      storage.masks[-1]).detach()
    return next_value


def reference_run(rollout_seed=99, act_seed=7, perm_seed=PERM_SEED):
    """Run the reference's own storage + PPO.update on a synthetic rollout; returns the fixture dict (inputs, returns in all
    four modes, normalised advantages, losses, per-parameter change digests)."""
    from . import ref_harness, ref_import

    ref_import.install_shims()
    from pytorchBaselines.a2c_ppo_acktr.algo.ppo import PPO
    from pytorchBaselines.a2c_ppo_acktr.model import Policy
    from pytorchBaselines.a2c_ppo_acktr.storage import SRNNRolloutStorage

    torch.set_num_threads(1)
    cfg = ref_harness.make_reference_config(**{"training.cuda": False, "training.num_processes": N})
    spaces = {"robot_node": ref_import.Box(-np.inf, np.inf, (1, 7)), "temporal_edges": ref_import.Box(-np.inf, np.inf, (1, 2)),
              "spatial_edges": ref_import.Box(-np.inf, np.inf, (H, 2))}
    act_space = ref_import.Box(-np.inf, np.inf, (2,))
    out = {}
    for tag, proper in (("gae", False), ("gae_proper", True), ("nogae", False), ("nogae_proper", True)):
        policy = Policy(spaces, act_space, base="srnn", base_kwargs=cfg)
        policy.load_state_dict(torch.load(os.path.join(ref_import.REFERENCE_ROOT, "data/example_model/checkpoints/27776.pt"),
                                          map_location="cpu"))
        storage = SRNNRolloutStorage(T, N, spaces, act_space, 128, 256, recurrent_cell_type="GRU")
        torch.manual_seed(act_seed)
        next_value = synthetic_rollout(policy, storage, seed=rollout_seed)
        storage.compute_returns(next_value, tag.startswith("gae"), cfg.reward.gamma, cfg.ppo.gae_lambda, proper)
        out["returns_" + tag] = storage.returns.numpy().copy()
        if tag != "gae":
            continue
        # ---- inputs (shared by all four return variants) ----
        for k in storage.obs:
            out["obs_" + k] = storage.obs[k].numpy().copy()
        out["h_node0"] = storage.recurrent_hidden_states["human_node_rnn"][0].numpy().copy()
        out["h_edge0"] = storage.recurrent_hidden_states["human_human_edge_rnn"][0].numpy().copy()
        for name in ("rewards", "value_preds", "action_log_probs", "actions", "masks", "bad_masks"):
            out[name] = getattr(storage, name).numpy().copy()
        out["next_value"] = next_value.numpy().copy()
        # ---- the update itself ----
        agent = PPO(policy, cfg.ppo.clip_param, cfg.ppo.epoch, cfg.ppo.num_mini_batch, cfg.ppo.value_loss_coef,
                    cfg.ppo.entropy_coef, lr=cfg.training.lr, eps=cfg.training.eps, max_grad_norm=cfg.training.max_grad_norm)
        before = {k: v.detach().clone() for k, v in policy.state_dict().items()}
        adv = storage.returns[:-1] - storage.value_preds[:-1]
        out["advantages"] = ((adv - adv.mean()) / (adv.std() + 1e-5)).numpy().copy()
        torch.manual_seed(perm_seed)
        losses = agent.update(storage)
        out["losses"] = np.asarray(losses, dtype=np.float64)
        names = sorted(before)
        out["param_names"] = np.asarray(names)
        for i, k in enumerate(names):
            after = policy.state_dict()[k].detach()
            delta = (after - before[k]).double().reshape(-1)
            out["delta_head_%02d" % i] = delta[:HEAD].numpy().copy()
            out["delta_stats_%02d" % i] = np.asarray([float(delta.norm()), float(delta.sum()), float(delta.abs().max())])
        out["hyper"] = np.asarray([cfg.ppo.clip_param, cfg.ppo.epoch, cfg.ppo.num_mini_batch, cfg.ppo.value_loss_coef,
                                   cfg.ppo.entropy_coef, cfg.training.lr, cfg.training.eps, cfg.training.max_grad_norm,
                                   cfg.reward.gamma, cfg.ppo.gae_lambda, perm_seed], dtype=np.float64)
    return out


def main():
    out = reference_run()
    print("losses", out["losses"])
    path = os.path.join(GOLDEN_DIR, "ppo_update_h5.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

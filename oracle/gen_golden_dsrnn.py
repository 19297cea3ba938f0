"""TEST INFRASTRUCTURE -- generates the DS-RNN forward fixtures by running the
REFERENCE's own `Policy.act(deterministic=True)` (pytorchBaselines/a2c_ppo_acktr/
model.py:63-86) on both shipped checkpoints.  Build container only:

    python -m oracle.gen_golden_dsrnn

Writes
  tests/golden/weights_holonomic_27776.npz  (data/example_model/checkpoints/27776.pt as float32 arrays)
  tests/golden/weights_unicycle_55554.npz   (data/example_model_unicycle/checkpoints/55554.pt)
  tests/golden/dsrnn_<ckpt>_h<H>.npz        inputs + reference outputs for H in {1,5,10,20}
and prints the torch restatement's (oracle/dsrnn_oracle.py) error against them.
"""
import os

import numpy as np
import torch

from . import dsrnn_oracle, ref_harness, ref_import

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CKPTS = {
    "holonomic_27776": "data/example_model/checkpoints/27776.pt",
    "unicycle_55554": "data/example_model_unicycle/checkpoints/55554.pt",
}


def sample_inputs(n, H, seed):
    g = torch.Generator().manual_seed(seed)
    robot_node = torch.cat([torch.rand(n, 1, 2, generator=g) * 12 - 6, torch.full((n, 1, 1), 0.3),
                            torch.rand(n, 1, 2, generator=g) * 12 - 6, torch.ones(n, 1, 1),
                            torch.rand(n, 1, 1, generator=g) * 6.28], -1)
    temporal = torch.randn(n, 1, 2, generator=g) * 0.6
    spatial = torch.randn(n, H, 2, generator=g) * 4.0
    spatial[::5, ::3] = 15.0 - robot_node[::5, :, :2]   # never-seen humans sit at (15,15) in the belief
    h_node = torch.randn(n, 1, 128, generator=g) * 0.5
    h_edge = torch.randn(n, H + 1, 256, generator=g) * 0.5
    masks = (torch.rand(n, 1, generator=g) > 0.2).float()
    return robot_node, temporal, spatial, h_node, h_edge, masks


def main():
    ref_import.install_shims()
    from pytorchBaselines.a2c_ppo_acktr.model import Policy

    torch.set_num_threads(4)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, rel in CKPTS.items():
        sd = torch.load(os.path.join(ref_import.REFERENCE_ROOT, rel), map_location="cpu")
        np.savez(os.path.join(GOLDEN_DIR, f"weights_{name}.npz"),
                 **{k: v.detach().cpu().numpy().astype(np.float32) for k, v in sd.items()})
        for H in (1, 5, 10, 20):
            n = 24
            cfg = ref_harness.make_reference_config(**{"sim.human_num": H, "training.cuda": False,
                                                       "training.num_processes": n})
            spaces = {"robot_node": ref_import.Box(-np.inf, np.inf, (1, 7)),
                      "temporal_edges": ref_import.Box(-np.inf, np.inf, (1, 2)),
                      "spatial_edges": ref_import.Box(-np.inf, np.inf, (H, 2))}
            act_space = ref_import.Box(-np.inf, np.inf, (2,))
            policy = Policy(spaces, act_space, base="srnn", base_kwargs=cfg)
            missing = policy.load_state_dict(sd)
            policy.eval()
            rn, te, se, hn, he, mk = sample_inputs(n, H, seed=1000 + H)
            hx = {"human_node_rnn": hn.clone(), "human_human_edge_rnn": he.clone()}
            with torch.no_grad():
                value, action, logp, hx_out = policy.act(
                    {"robot_node": rn, "temporal_edges": te, "spatial_edges": se}, hx, mk, deterministic=True)
                feat = policy.base({"robot_node": rn, "temporal_edges": te, "spatial_edges": se},
                                   {"human_node_rnn": hn.clone(), "human_human_edge_rnn": he.clone()}, mk, infer=True)[1]
            ref = dict(value=value, action_mean=action, log_prob=logp, actor_features=feat,
                       h_node=hx_out["human_node_rnn"], h_edge=hx_out["human_human_edge_rnn"])
            mine = dsrnn_oracle.forward(sd, rn, te, se, hn, he, mk)
            errs = {k: float((mine[k] - ref[k].reshape(mine[k].shape)).abs().max()) for k in
                    ("value", "action_mean", "actor_features", "h_node", "h_edge")}
            print(f"[{name} H={H}] {missing}; restatement max abs err: {errs}")
            np.savez_compressed(
                os.path.join(GOLDEN_DIR, f"dsrnn_{name}_h{H}.npz"),
                robot_node=rn.numpy(), temporal_edges=te.numpy(), spatial_edges=se.numpy(), h_node=hn.numpy(),
                h_edge=he.numpy(), masks=mk.numpy(), **{"ref_" + k: v.numpy() for k, v in ref.items()})


if __name__ == "__main__":
    main()

"""cProfile of the host side of Policy.act + CrowdVecEnv.step at a tiny batch (the GPU work is negligible: what is left is the
Python / ctypes / launch cost that sits in the e2e loop's critical path after every host synchronisation). Development aid."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crowdnav_dsrnn_b200.envs import CrowdVecEnv  # noqa: E402
from crowdnav_dsrnn_b200.model import Policy  # noqa: E402
from crowdnav_dsrnn_b200.spaces import crowd_spaces  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c1"]
    N, H = wl["envs_per_gpu"], wl["human_num"]
    dev = torch.device("cuda:0")
    cfg = bench.make_config(wl)
    venv = CrowdVecEnv(cfg, N, dev, seed=0, phase="train")
    obs_space, act_space = crowd_spaces(H)
    policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg)
    policy.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(wl["weights"]).items()})
    policy = policy.to(dev)
    obs = venv.reset()
    hx = {"human_node_rnn": torch.zeros(N, 1, 128, device=dev), "human_human_edge_rnn": torch.zeros(N, H + 1, 256, device=dev)}
    masks = torch.zeros(N, 1, device=dev)
    pin_action = torch.empty(N, 2, dtype=torch.float32).pin_memory()
    pin_masks = torch.empty(N, 1, dtype=torch.float32).pin_memory()

    def loop(steps):
        nonlocal obs, hx, masks
        for _ in range(steps):
            _, action, _, hx = policy.act(obs, hx, masks, deterministic=True)
            pin_action.copy_(action, non_blocking=True)
            obs, reward, done, infos = venv.step(pin_action.to(dev, non_blocking=True))
            pin_masks.copy_(torch.from_numpy(1.0 - done.astype(np.float32)).unsqueeze(1))
            masks = pin_masks.to(dev, non_blocking=True)

    loop(50)
    pr = cProfile.Profile()
    pr.enable()
    loop(300)
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()

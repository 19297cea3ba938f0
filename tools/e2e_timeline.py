"""Host-side timeline of the e2e loop of bench.py (reference-style: act -> step with host buffers -> masks): where the CPU
time between the host synchronisation of one step and the first kernel of the next goes. Development aid."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crowdnav_dsrnn_b200.envs import CrowdVecEnv  # noqa: E402
from crowdnav_dsrnn_b200.model import Policy  # noqa: E402
from crowdnav_dsrnn_b200.spaces import crowd_spaces  # noqa: E402


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
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
    acc = np.zeros(5)
    steps = 60
    for it in range(steps + 10):
        t0 = time.perf_counter()
        _, action, _, hx = policy.act(obs, hx, masks, deterministic=True)
        t1 = time.perf_counter()
        pin_action.copy_(action, non_blocking=True)
        a_dev = pin_action.to(dev, non_blocking=True)
        t2 = time.perf_counter()
        obs, reward, done, infos = venv.step(a_dev)
        t3 = time.perf_counter()
        pin_masks.copy_(torch.from_numpy(1.0 - done.astype(np.float32)).unsqueeze(1))
        masks = pin_masks.to(dev, non_blocking=True)
        t4 = time.perf_counter()
        if it >= 10:
            acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0]
    acc *= 1e6 / steps
    print("per step (us): act launch %.0f | action copies %.0f | env.step (launch + D2H + sync wait) %.0f | masks %.0f | total %.0f"
          % tuple(acc))


if __name__ == "__main__":
    main()

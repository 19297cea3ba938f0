"""Device-resident rollout at the c3 shape: plain graph replay (rollout.GraphedRollout) vs the two-half software pipeline
(rollout.PipelinedRollout), a few split points.  usage: python tools/bench_pipeline.py [n_envs] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from crowdnav_dsrnn_b200.envs import CrowdVecEnv  # noqa: E402
from crowdnav_dsrnn_b200.model import Policy  # noqa: E402
from crowdnav_dsrnn_b200.rollout import GraphedRollout, PipelinedRollout  # noqa: E402
from crowdnav_dsrnn_b200.spaces import crowd_spaces  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    wl = bench.WORKLOADS[sys.argv[3] if len(sys.argv) > 3 else "c3"]
    dev = torch.device("cuda:0")
    cfg = bench.make_config(wl)
    H = wl["human_num"]
    cfg.training.num_processes = n
    obs_space, act_space = crowd_spaces(H)
    policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg)
    policy.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(wl["weights"]).items()})
    policy = policy.to(dev)

    def timed(fn, prime=150):
        for _ in range(prime):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    venv = CrowdVecEnv(cfg, n, dev, seed=0, phase="train")
    roll = GraphedRollout(policy, venv, venv.reset())
    ms = timed(roll.step)
    print("plain      %6d envs  %.4f ms/step  %.2f M env-steps/s" % (n, ms, n / ms / 1e3), flush=True)
    roll.close()
    venv.close()
    d = PipelinedRollout.default_split(n, H, dev)
    for split in (d, n // 2, d - 256, d + 148):
        if not 0 < split < n:
            continue
        pr = PipelinedRollout(policy, cfg, n, dev, seed=0, phase="train", split=split)
        ms = timed(pr.step)
        print("pipelined  %6d envs  split %6d  %.4f ms/step  %.2f M env-steps/s" % (n, split, ms, n / ms / 1e3), flush=True)
        pr.close()


if __name__ == "__main__":
    main()

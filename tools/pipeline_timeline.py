"""Kernel timeline (CUPTI through torch.profiler) of a few replays of rollout.PipelinedRollout / GraphedRollout.
usage: python tools/pipeline_timeline.py [plain|pipe] [n_envs] [split] [workload]"""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from crowdnav_dsrnn_b200.envs import CrowdVecEnv  # noqa: E402
from crowdnav_dsrnn_b200.model import Policy  # noqa: E402
from crowdnav_dsrnn_b200.rollout import GraphedRollout, PipelinedRollout  # noqa: E402
from crowdnav_dsrnn_b200.spaces import crowd_spaces  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "pipe"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    split = int(sys.argv[3]) if len(sys.argv) > 3 and int(sys.argv[3]) > 0 else None
    wl = bench.WORKLOADS[sys.argv[4] if len(sys.argv) > 4 else "c3"]
    dev = torch.device("cuda:0")
    cfg = bench.make_config(wl)
    H = wl["human_num"]
    cfg.training.num_processes = n
    obs_space, act_space = crowd_spaces(H)
    policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg)
    policy.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(wl["weights"]).items()})
    policy = policy.to(dev)
    if mode == "plain":
        venv = CrowdVecEnv(cfg, n, dev, seed=0, phase="train")
        img = os.environ.get("CN_IMG_MODE", "both")          # both | off | in | out (in / out: timing only, stale images)
        roll = GraphedRollout(policy, venv, venv.reset(), edge_image={"both": True, "off": False}.get(img, img))
    else:
        roll = PipelinedRollout(policy, cfg, n, dev, seed=0, phase="train", split=split)
    for _ in range(150):
        roll.step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for _ in range(4):
            roll.step()
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    for e in ev:
        print("%9.1f us  +%7.1f us  stream %3s  %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:60]))


if __name__ == "__main__":
    main()

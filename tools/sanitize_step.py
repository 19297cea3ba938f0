"""Small crowd-step / reset workload for compute-sanitizer (memcheck, racecheck, initcheck): every step-kernel
instantiation the launcher can pick (group sizes 4..32, the optional-behaviour build, the two-launch form, the thread-per-human
form) and both reset kernels, with auto-reset on.

    compute-sanitizer --tool racecheck python tools/sanitize_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crowdnav_dsrnn_b200 import Config  # noqa: E402
from crowdnav_dsrnn_b200.engine import CrowdEngine  # noqa: E402


def run(n, steps, human_num, kinematics="holonomic", env=None, **over):
    for k, v in (env or {}).items():
        os.environ[k] = v
    cfg = Config(kinematics=kinematics, human_num=human_num)
    for k, v in over.items():
        sec, _, attr = k.partition(".")
        setattr(getattr(cfg, sec), attr, v)
    eng = CrowdEngine(cfg, n, torch.device("cuda:0"), phase="train", seed=1)
    eng.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(steps):
        eng.step(torch.randn(n, 2, device="cuda", generator=g) * 0.7, auto_reset=True)
    torch.cuda.synchronize()
    eng.close()
    for k in (env or {}):
        os.environ.pop(k, None)
    print("ok", human_num, kinematics, env or "", over, flush=True)


def main():
    run(96, 12, 3)                                             # G = 4
    run(96, 12, 7, **{"robot.visible": True})                  # G = 8
    run(64, 12, 12, kinematics="unicycle")                     # G = 16
    run(64, 12, 20, **{"robot.FOV": 0.5})                      # G = 32
    run(64, 12, 20, env={"CN_STEP_SPLIT": "1"})                # two launches
    run(64, 12, 20, env={"CN_STEP_SEQ": "1"})                  # thread per human
    run(64, 12, 9, **{"humans.random_policy_changing": True, "humans.random_unobservability": True,
                      "humans.random_radii": True, "humans.random_v_pref": True, "humans.FOV": 1.5})
    run(64, 12, 13, **{"sim.group_human": True})


if __name__ == "__main__":
    main()

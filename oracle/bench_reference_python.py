"""TEST INFRASTRUCTURE -- the CPU baseline BASELINE.md section 3 / SURVEY.md 8(d) specify: the reference's OWN Python env
(crowd_sim/, crowd_nav/policy/) under its OWN `ShmemVecEnv` (pytorchBaselines/a2c_ppo_acktr/shmem_vec_env.py, fork context,
envs.py:136-139) and `VecPyTorch`, stepped with its OWN `Policy.act` exactly as train.py:226-261 does -- unmodified, imported
from /root/reference under oracle/ref_import.py shims (gym / baselines / matplotlib stand-ins, exact-geometry shapely,
rvo2 = the restated RVO2 of oracle/rvo2_port.c).  "reference-equivalent CPU path (rvo2 restated)".

The reference is Python and cannot travel to the GPU box, so this runs in the BUILD CONTAINER only:

    python -m oracle.bench_reference_python [--seconds 30] > profiles/r2_reference_python_cpu.json

One worker process per host core (N_cpu = os.cpu_count()), one env per process, lock-step; env-steps/s = N_cpu * steps /
elapsed with the DS-RNN forward included; workloads C1 (H=5 holonomic), C2 (H=10 unicycle, dt 0.1), C3 (H=20, FOV 0.5 pi,
wall-capped: the unbounded goal re-sampling loop of crowd_sim.py:787-806 dominates there)."""
import argparse
import json
import os
import platform
import time

import numpy as np

WORKLOADS = {
    "c1": dict(over={"sim.train_val_sim": ["circle_crossing"]}, ckpt="data/example_model/checkpoints/27776.pt"),
    "c2": dict(over={"action_space.kinematics": "unicycle", "sim.human_num": 10, "env.time_step": 0.1,
                     "reward.discomfort_penalty_factor": 1.0}, ckpt="data/example_model_unicycle/checkpoints/55554.pt"),
    "c3": dict(over={"sim.human_num": 20, "robot.FOV": 0.5}, ckpt="data/example_model/checkpoints/27776.pt"),
}


def run(name, seconds, warmup_steps, n_cpu):
    import torch

    from . import ref_harness, ref_import

    ref_import.install_shims()
    import crowd_sim  # noqa: F401  (registers the gym ids)
    from pytorchBaselines.a2c_ppo_acktr.envs import make_vec_envs
    from pytorchBaselines.a2c_ppo_acktr.model import Policy

    wl = WORKLOADS[name]
    over = dict(wl["over"])
    over.update({"training.cuda": False, "training.num_processes": n_cpu, "env.seed": 0})
    cfg = ref_harness.make_reference_config(**over)
    unicycle = cfg.action_space.kinematics == "unicycle"
    if unicycle:
        # declared oracle patch P1: crowd_sim.py:1004-1005,1023 read .vx/.vy from an ActionRot and crash; the
        # worker processes are forked from this one, so the patched constructor is what they see
        from crowd_sim.envs import crowd_sim_dict as csd
        from crowd_sim.envs.utils.action import ActionRot

        class _ActionRotP1(ActionRot):
            pass

        def make(v, r):
            a = _ActionRotP1(v, r)
            a.vx, a.vy = float(v), 0.0          # only the SM4/SM5 info fields read them (speed / jerk of the commanded velocity)
            return a

        csd.ActionRot = make
    device = torch.device("cpu")
    torch.set_num_threads(max(1, n_cpu // 2))
    envs = make_vec_envs(cfg.env.env_name, cfg.env.seed, n_cpu, cfg.reward.gamma, None, device, False, config=cfg)
    policy = Policy(envs.observation_space.spaces, envs.action_space, base="srnn", base_kwargs=cfg)
    policy.load_state_dict(torch.load(os.path.join(ref_import.REFERENCE_ROOT, wl["ckpt"]), map_location="cpu"))
    H = cfg.sim.human_num
    hx = {"human_node_rnn": torch.zeros(n_cpu, 1, 128), "human_human_edge_rnn": torch.zeros(n_cpu, H + 1, 256)}
    masks = torch.zeros(n_cpu, 1)
    obs = envs.reset()
    steps, t0, t_start, capped = 0, None, time.perf_counter(), False
    while True:
        if steps == warmup_steps:
            t0 = time.perf_counter()
        with torch.no_grad():
            _, action, _, hx = policy.act(obs, hx, masks, deterministic=True)
        obs, reward, done, infos = envs.step(action)
        masks = torch.FloatTensor([[0.0] if d else [1.0] for d in done])
        steps += 1
        now = time.perf_counter()
        if t0 is not None and now - t0 >= seconds:
            break
        if t0 is None and now - t_start >= 4 * seconds:      # H=20: the warm-up itself may not finish; time what there is
            t0, warmup_steps, capped = t_start, 0, True
            break
    elapsed = time.perf_counter() - t0
    timed = steps - warmup_steps
    envs.close()
    return {"workload": name, "human_num": H, "kinematics": cfg.action_space.kinematics, "processes": n_cpu,
            "timed_steps": timed, "warmup_steps": warmup_steps, "elapsed_s": elapsed, "wall_capped": capped,
            "env_steps_per_s": n_cpu * timed / elapsed}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--warmup-steps", type=int, default=200)
    ap.add_argument("--workloads", default="c1,c2,c3")
    ap.add_argument("--out", default=None, help="JSON file, rewritten after every workload")
    args = ap.parse_args()
    from . import crowd_oracle

    crowd_oracle.build()
    n_cpu = os.cpu_count() or 1
    rows = []
    for w in args.workloads.split(","):
        rows.append(run(w, args.seconds, args.warmup_steps, n_cpu))
        doc = {"what": "reference-equivalent CPU path (reference Python env + ShmemVecEnv + Policy.act, rvo2 restated)",
               "where": "build container, %d vCPU (%s)" % (n_cpu, platform.processor() or platform.machine()),
               "kind": "reference-python", "results": rows}
        if args.out:                      # written after every workload: the H=20 one may have to be killed from outside
            with open(args.out, "w") as f:
                json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()

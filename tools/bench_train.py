"""Training throughput probe (BASELINE.json configs[4] shape): rollout (CUDA) + PPO update (torch autograd) per update.

    python tools/bench_train.py --envs 4096 --humans 20 --updates 3 --per-pass 1024
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_train.py ...

Prints one JSON line: env-steps/s over whole updates (rollout + update), and the split between the two.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from crowdnav_dsrnn_b200 import Config  # noqa: E402
from crowdnav_dsrnn_b200 import train as train_mod  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--humans", type=int, default=20)
    ap.add_argument("--updates", type=int, default=3)
    ap.add_argument("--per-pass", type=int, default=1024)
    ap.add_argument("--tf32", action="store_true", help="TF32 tensor cores for the torch GEMMs of the update (fwd + bwd)")
    ap.add_argument("--bf16x3", action="store_true", help="split-bf16 3-pass tensor-core GEMMs for the recurrent products of the update")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if args.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = Config()
    cfg.sim.human_num = args.humans
    cfg.training.num_processes = args.envs
    cfg.training.log_interval = 1
    dev = torch.device("cuda", local)
    # time the update separately by wrapping PPO.update
    from crowdnav_dsrnn_b200 import ppo as ppo_mod
    spent = {"update": 0.0}
    orig = ppo_mod.PPO.update

    def timed(self, rollouts):
        torch.cuda.synchronize()
        t0 = time.time()
        out = orig(self, rollouts)
        torch.cuda.synchronize()
        spent["update"] += time.time() - t0
        return out

    ppo_mod.PPO.update = timed
    train_mod.train(cfg, dev, num_updates=1, output_dir=None, log=None, max_envs_per_pass=args.per_pass, bf16x3_update=args.bf16x3)   # warm-up
    spent["update"] = 0.0
    torch.cuda.synchronize()
    t0 = time.time()
    _, hist = train_mod.train(cfg, dev, num_updates=args.updates, output_dir=None, log=None, max_envs_per_pass=args.per_pass,
                               bf16x3_update=args.bf16x3)
    torch.cuda.synchronize()
    wall = time.time() - t0
    steps = args.updates * args.envs * world * cfg.ppo.num_steps
    # train() starts its own clock after building the envs, the policy and the storage (0.03-0.3 s of host work that is not
    # part of the steady state): use it for the throughput, keep the outer clock as `wall_incl_setup_s`
    loop = steps / max(1, hist[-1]["fps"])
    if local == 0:
        print(json.dumps({"metric": "training env-steps/s (rollout + PPO update)", "value": steps / loop, "n_gpus": world,
                          "envs_per_gpu": args.envs, "humans": args.humans, "updates": args.updates,
                          "update_seconds_per_update": spent["update"] / args.updates,
                          "rollout_seconds_per_update": (loop - spent["update"]) / args.updates, "wall_incl_setup_s": wall,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                          "last": {k: hist[-1][k] for k in ("loss/value_loss", "loss/policy_loss", "success", "episodes")}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU, world_size = 2 over gloo: envs shard by global id with no data-path collective; a 2-rank rollout of the
shards (env_id_offset = rank * N/2, nenv = N) equals the unsharded rollout, i.e. results do not depend on the
number of GPUs.  The compute here is the oracle (CPU); what is under test is the sharding contract
that bench.py / CrowdVecEnv use: flatten_config(..., env_id_offset, nenv) + max-over-ranks timing reduction."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crowdnav_dsrnn_b200 import Config, abi
from oracle import crowd_oracle

N, H, STEPS = 32, 5, 25


def _rollout(cfg, n, actions):
    st = crowd_oracle.OracleState(n, H)
    crowd_oracle.reset(cfg, st)
    dones = 0
    for a in actions:
        out = crowd_oracle.step(cfg, st, a, auto_reset=True)
        dones += int(out.done.sum())
    return st, dones


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg_obj = Config(human_num=H)
    per = N // world
    cfg = abi.flatten_config(cfg_obj, per, phase="train", env_id_offset=rank * per, nenv=N)
    rng = np.random.default_rng(0)
    actions = [rng.normal(0, 0.7, (N, 2)).astype(np.float32) for _ in range(STEPS)]
    st, dones = _rollout(cfg, per, [a[rank * per:(rank + 1) * per] for a in actions])
    humans = [torch.zeros(per, H, 9) for _ in range(world)]
    dist.all_gather(humans, torch.from_numpy(st.humans))
    t = torch.tensor([float(rank + 1), float(dones)])            # bench.py: MAX over ranks for time, SUM for counters
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        ret["humans"] = torch.cat(humans).numpy()
        ret["tmax"], ret["dones"] = float(tmax[0]), int(tsum[1])
    dist.destroy_process_group()


def test_two_rank_sharded_rollout_equals_single_process():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29533, ret), nprocs=2, join=True)
    cfg = abi.flatten_config(Config(human_num=H), N, phase="train", env_id_offset=0, nenv=N)
    rng = np.random.default_rng(0)
    actions = [rng.normal(0, 0.7, (N, 2)).astype(np.float32) for _ in range(STEPS)]
    st, dones = _rollout(cfg, N, actions)
    assert np.array_equal(ret["humans"], st.humans)
    assert ret["dones"] == dones and ret["tmax"] == 2.0

"""Per-kernel timing on one GPU (CUDA events): DS-RNN forward (and its edge stage) and the crowd step kernel.
Development aid; the contract numbers come from bench.py."""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crowdnav_dsrnn_b200 import _lib  # noqa: E402
from crowdnav_dsrnn_b200.envs import CrowdVecEnv  # noqa: E402
from crowdnav_dsrnn_b200.model import Policy  # noqa: E402
from crowdnav_dsrnn_b200.spaces import crowd_spaces  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--precisions", default="bf16x3,bf16,fp32")
    args = ap.parse_args()
    wl = bench.WORKLOADS[args.workload]
    N, H = args.envs or wl["envs_per_gpu"], wl["human_num"]
    dev = torch.device("cuda:0")
    lib = _lib.load()
    cfg = bench.make_config(wl)
    venv = CrowdVecEnv(cfg, N, dev, seed=0, phase="train")
    obs_space, act_space = crowd_spaces(H)
    policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg)
    policy.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(wl["weights"]).items()})
    policy = policy.to(dev)
    obs = venv.reset()
    hx = {"human_node_rnn": torch.randn(N, 1, 128, device=dev) * 0.3,
          "human_human_edge_rnn": torch.randn(N, H + 1, 256, device=dev) * 0.3}
    masks = torch.ones(N, 1, device=dev)
    for prec in args.precisions.split(","):
        policy.precision = prec
        for _ in range(3):
            policy.act(obs, dict(hx), masks, deterministic=True)
        lib.cn_dsrnn_enable_timing(policy._handle, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.iters):
            policy.act(obs, dict(hx), masks, deterministic=True)
        e1.record()
        torch.cuda.synchronize()
        ms, n = C.c_float(), C.c_int()
        lib.cn_dsrnn_time_ms(policy._handle, C.byref(ms), C.byref(n))
        lib.cn_dsrnn_enable_timing(policy._handle, 0)
        fw = e0.elapsed_time(e1) / args.iters
        edge = ms.value / max(1, n.value)
        tf = N * bench.edge_stage_flops(H) / (edge * 1e-3) / 1e12
        print("forward[%s] N=%d H=%d: %.3f ms/forward, edge stage %.3f ms (%.1f algorithmic TFLOP/s), rest %.3f ms"
              % (prec, N, H, fw, edge, tf, fw - edge))
    lib.cn_env_enable_timing(venv.engine.handle, 1)
    act = torch.randn(N, 2, device=dev) * 0.5
    for _ in range(100):
        venv.step_device(act)
    lib.cn_env_time_ms(venv.engine.handle, C.byref(C.c_float()), C.byref(C.c_int()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        venv.step_device(act)
    e1.record()
    torch.cuda.synchronize()
    ms, n = C.c_float(), C.c_int()
    lib.cn_env_time_ms(venv.engine.handle, C.byref(ms), C.byref(n))
    st = ms.value / max(1, n.value)
    print("env step N=%d H=%d: %.3f ms step+reset, step kernel %.3f ms (%.1f GB/s algorithmic)"
          % (N, H, e0.elapsed_time(e1) / args.iters, st, N * bench.step_bytes(H) / (st * 1e-3) / 1e9))


if __name__ == "__main__":
    main()

"""Where the PPO update's time goes: one forward + backward pass of `Policy.evaluate_actions`
over a [T, n] rollout chunk under torch.profiler, kernels grouped by name.

    python tools/profile_update.py --envs 1024 --humans 20 --steps 30 [--tf32] > gpurun_out/update_profile.txt
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from crowdnav_dsrnn_b200 import Config  # noqa: E402
from crowdnav_dsrnn_b200.model import Policy  # noqa: E402
from crowdnav_dsrnn_b200.spaces import crowd_spaces  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--humans", type=int, default=20)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--tf32", action="store_true")
    ap.add_argument("--bf16x3", action="store_true", help="model.SEQUENCE_GEMM = 'bf16x3'")
    ap.add_argument("--per-step", action="store_true", help="the per-step autograd graph (Policy.sequence_impl='per_step')")
    ap.add_argument("--native", action="store_true", help="Policy.sequence_impl='native': every contraction on the library's tcgen05 kernels")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = args.tf32
    if args.bf16x3:
        from crowdnav_dsrnn_b200 import model as model_mod
        model_mod.SEQUENCE_GEMM = "bf16x3"
    n, H, T = args.envs, args.humans, args.steps
    cfg = Config(human_num=H)
    obs_space, act_space = crowd_spaces(H)
    policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg).to(dev)
    if args.per_step:
        policy.sequence_impl = "per_step"
    if args.native:
        policy.sequence_impl = "native"
    g = torch.Generator(device=dev).manual_seed(0)
    obs = {"robot_node": torch.randn(T * n, 1, 7, device=dev, generator=g),
           "temporal_edges": torch.randn(T * n, 1, 2, device=dev, generator=g),
           "spatial_edges": torch.randn(T * n, H, 2, device=dev, generator=g) * 3}
    masks = (torch.rand(T * n, 1, device=dev, generator=g) > 0.03).float()
    action = torch.randn(T * n, 2, device=dev, generator=g)

    def one_pass():
        hx = {"human_node_rnn": torch.randn(n, 1, 128, device=dev, generator=g) * 0.3,
              "human_human_edge_rnn": torch.randn(n, H + 1, 256, device=dev, generator=g) * 0.3}
        policy.zero_grad(set_to_none=True)
        value, logp, ent, _ = policy.evaluate_actions(obs, hx, masks, action)
        (value.mean() + logp.mean() - 0.01 * ent).backward()

    for _ in range(2):
        one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        one_pass()
    e1.record()
    torch.cuda.synchronize()
    print("pass of %d envs x %d humans x %d steps: %.2f ms (tf32=%s, per_step=%s, bf16x3=%s, native=%s), peak mem %.2f GB" % (
        n, H, T, e0.elapsed_time(e1) / 3, args.tf32, args.per_step, args.bf16x3, args.native, torch.cuda.max_memory_allocated() / 2**30))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        one_pass()
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type is not None and str(ev.device_type).endswith("CUDA"):
            name = ev.name[:90]
            t, c = rows.get(name, (0.0, 0))
            rows[name] = (t + ev.device_time, c + 1)
    total = sum(t for t, _ in rows.values())
    print("GPU kernel time %.2f ms in %d launches" % (total / 1e3, sum(c for _, c in rows.values())))
    for name, (t, c) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:30]:
        print("%8.2f ms %5.1f%% %6d  %s" % (t / 1e3, 100 * t / total, c, name))


if __name__ == "__main__":
    main()

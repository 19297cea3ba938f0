"""cn_gemm_bf16x3 on the shapes of the PPO update (CUDA events, L2 flushed between timed launches by the operand sizes).
    python tools/bench_gemm.py [--envs 4096] [--humans 20] [--steps 30]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crowdnav_dsrnn_b200 import native  # noqa: E402

DEV = torch.device("cuda:0")


def pair(*shape):
    return (torch.randn(*shape, device=DEV).bfloat16(), (torch.randn(*shape, device=DEV) * 1e-3).bfloat16())


def timeit(name, fn, flops, bytes_, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%-58s %8.3f ms  %7.1f TFLOP/s algorithmic (x3 issued)  %6.2f TB/s operand bytes" % (name, ms, flops / ms / 1e9, bytes_ / ms / 1e9), flush=True)
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--humans", type=int, default=20)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    n, H, T = args.envs, args.humans, args.steps
    S = n * H
    # per-step recurrent product of the backward: d += G[:, 256:] W_hh   (spatial + temporal in one launch)
    g_s, g_t = pair(S, 1024), pair(n, 1024)
    whh = pair(768, 256)
    d_s, d_t = torch.zeros(S, 256, device=DEV), torch.zeros(n, 256, device=DEV)
    rows = S + n
    timeit("bwd recurrent [%d+%d, 768] x [768, 256] accumulate" % (S, n),
           lambda: native.gemm([dict(a=(g_s[0][:, 256:], g_s[1][:, 256:]), b=whh, b_mn=True, c=d_s, accumulate=True),
                                dict(a=(g_t[0][:, 256:], g_t[1][:, 256:]), b=whh, b_mn=True, c=d_t, accumulate=True)]),
           2.0 * rows * 768 * 256, rows * (768 * 4 + 256 * 8))
    # whole-sequence products on T' steps worth of rows (T' chosen to stay within memory)
    Tq = min(T, max(1, (6 << 30) // (S * 1024 * 4)))
    R = Tq * S
    G = pair(R, 1024)
    hm, e = pair(R, 256), pair(R, 64)
    wih = pair(768, 64)
    de = torch.empty(R, 64, device=DEV)
    timeit("dx: [%d, 768] x [768, 64]" % R, lambda: native.gemm([dict(a=(G[0][:, :768], G[1][:, :768]), b=wih, b_mn=True, c=de)]),
           2.0 * R * 768 * 64, R * (768 * 4 + 64 * 4))
    dw = torch.zeros(768, 320, device=DEV)
    timeit("dW_hh: [%d, 768]^T x [%d, 256] split-K" % (R, R),
           lambda: native.gemm([dict(a=(G[0][:, 256:], G[1][:, 256:]), a_mn=True, b=hm, b_mn=True, c=dw[:, :256], split_k=0)]),
           2.0 * R * 768 * 256, R * (768 * 4 + 256 * 4))
    timeit("dW_ih: [%d, 768]^T x [%d, 64] split-K" % (R, R),
           lambda: native.gemm([dict(a=(G[0][:, :768], G[1][:, :768]), a_mn=True, b=e, b_mn=True, c=dw[:, 256:], split_k=0)]),
           2.0 * R * 768 * 64, R * (768 * 4 + 64 * 4))
    timeit("dW_hh + dW_ih grouped",
           lambda: native.gemm([dict(a=(G[0][:, 256:], G[1][:, 256:]), a_mn=True, b=hm, b_mn=True, c=dw[:, :256], split_k=0),
                                dict(a=(G[0][:, :768], G[1][:, :768]), a_mn=True, b=e, b_mn=True, c=dw[:, 256:], split_k=0)]),
           2.0 * R * 768 * 320, R * (768 * 8 + 320 * 4))
    del G, hm, e, de
    # the batched linears over the T*n samples
    M = T * n
    for name, k, nn_ in (("edge_attention_embed", 512, 64), ("actor.0", 256, 256), ("node gi", 128, 384), ("output_linear", 128, 256), ("attn q", 256, 64)):
        x, w = pair(M, k), pair(nn_, k)
        y = torch.empty(M, nn_, device=DEV)
        timeit("linear %s: [%d, %d] x [%d, %d]^T" % (name, M, k, nn_, k), lambda: native.gemm([dict(a=x, b=w, c=y)]),
               2.0 * M * k * nn_, M * (k * 4 + nn_ * 4))
        dyp = pair(M, nn_)
        dwl = torch.zeros(nn_, k, device=DEV)
        timeit("   its dW: [%d, %d]^T x [%d, %d] split-K" % (M, nn_, M, k), lambda: native.gemm([dict(a=dyp, a_mn=True, b=x, b_mn=True, c=dwl, split_k=0)]),
               2.0 * M * k * nn_, M * (k * 4 + nn_ * 4))
    # node GRU per-step product (tiny M)
    x, w = pair(n, 128), pair(384, 128)
    y = torch.empty(n, 384, device=DEV)
    timeit("node gh: [%d, 128] x [384, 128]^T" % n, lambda: native.gemm([dict(a=x, b=w, c=y)]), 2.0 * n * 128 * 384, n * (128 * 4 + 384 * 4), iters=30)


if __name__ == "__main__":
    main()

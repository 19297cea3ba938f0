"""Development aid: cn_gemm_bf16x3 on a few shapes, printing the error and, on a mismatch, where it sits (per 32 x 32 block)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crowdnav_dsrnn_b200 import native  # noqa: E402

DEV = torch.device("cuda:0")


def rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def report(name, got, want):
    err = (got.double() - want).abs()
    scale = max(1.0, want.abs().max().item())
    bad = err > 1e-4 * scale
    print("%-40s max err %.3e (scale %.2f) bad %d / %d nan %d" % (name, err.max().item(), scale, int(bad.sum()), bad.numel(),
                                                                   int(torch.isnan(got).sum())), flush=True)
    if bad.any():
        m, n = err.shape
        rb, cb = (m + 31) // 32, (n + 31) // 32
        for i in range(min(rb, 8)):
            print("   rows %4d+: " % (i * 32) + " ".join("%8.1e" % err[i * 32:(i + 1) * 32, j * 32:(j + 1) * 32].max().item() for j in range(min(cb, 10))))
        r, c = [int(v) for v in torch.nonzero(bad)[0]]
        print("   first bad (%d, %d): got %g want %g" % (r, c, got[r, c].item(), want[r, c].item()), flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "kk"):
        for m, n, k in ((128, 64, 64), (128, 256, 64), (128, 256, 128), (300, 80, 200), (1000, 320, 768)):
            x, w = rand(m, k, seed=1), rand(n, k, seed=2, scale=k ** -0.5)
            y = torch.full((m, n), float("nan"), device=DEV)
            native.gemm([dict(a=native.split(x), b=native.split(w), c=y)])
            torch.cuda.synchronize()
            report("K-major x K-major %dx%dx%d" % (m, n, k), y, x.double() @ w.double().t())
    if which in ("all", "kmn"):
        for m, n, k in ((128, 64, 64), (128, 256, 128), (256, 256, 768), (77, 64, 768), (130, 200, 72)):
            x, w = rand(m, k, seed=3), rand(k, n, seed=4, scale=k ** -0.5)
            y = torch.full((m, n), float("nan"), device=DEV)
            native.gemm([dict(a=native.split(x), b=native.split(w), b_mn=True, c=y)])
            torch.cuda.synchronize()
            report("K-major x MN-major %dx%dx%d" % (m, n, k), y, x.double() @ w.double())
    if which in ("all", "mnmn"):
        for rows, m, n, split in ((64, 128, 64, 1), (128, 128, 256, 1), (1000, 768, 256, 1), (5000, 768, 256, 0), (70000, 256, 512, 0)):
            dy, x = rand(rows, m, seed=5), rand(rows, n, seed=6)
            c = torch.zeros(m, n, device=DEV)
            native.gemm([dict(a=native.split(dy), a_mn=True, b=native.split(x), b_mn=True, c=c, split_k=split)])
            torch.cuda.synchronize()
            report("MN-major x MN-major rows %d -> %dx%d split %d" % (rows, m, n, split), c, dy.double().t() @ x.double())
    if which == "mc":           # multicast debugging: B MN-major, two row tiles, various atom / k-block counts
        for m, n, k in ((256, 64, 64), (256, 128, 64), (256, 256, 64), (256, 256, 128), (256, 192, 64), (512, 256, 64)):
            x = rand(m, k, seed=3)
            w = torch.arange(k, device=DEV, dtype=torch.float32).view(k, 1) * 0 + torch.arange(n, device=DEV, dtype=torch.float32).view(1, n) / 16 + 1
            w = w + torch.arange(k, device=DEV, dtype=torch.float32).view(k, 1) / 64
            y = torch.full((m, n), float("nan"), device=DEV)
            native.gemm([dict(a=native.split(x), b=native.split(w), b_mn=True, c=y)])
            torch.cuda.synchronize()
            report("mc K-major x MN-major %dx%dx%d" % (m, n, k), y, x.double() @ w.double())
            # which B element would explain row 128? solve per column with one-hot A rows
            eye = torch.zeros(m, k, device=DEV)
            eye[128:128 + min(k, 128), :] = torch.eye(k, device=DEV)[:min(k, 128)]
            y2 = torch.full((m, n), float("nan"), device=DEV)
            native.gemm([dict(a=native.split(eye), b=native.split(w), b_mn=True, c=y2)])
            torch.cuda.synchronize()
            got = y2[128:128 + min(k, 128)]            # should equal w[:k]
            bad = (got - w[:got.shape[0]]).abs() > 1e-3
            print("   one-hot probe: %d bad of %d; bad k rows %s ; bad n cols %s" % (int(bad.sum()), bad.numel(),
                  sorted(set(torch.nonzero(bad)[:, 0].tolist()))[:40], sorted(set((torch.nonzero(bad)[:, 1] // 16).tolist()))), flush=True)
            if bad.any():
                r, c = [int(v) for v in torch.nonzero(bad)[0]]
                print("   e.g. B[k=%d, n=%d]: got %g want %g" % (r, c, got[r, c].item(), w[r, c].item()), flush=True)
    if which in ("all", "ext"):       # the G^T [e | 1] product of the edge-GRU backward: M = 1024, N = 72 (ld 72)
        for rows in (42, 210, 5000):
            gfull, e = rand(rows, 1024, seed=9), rand(rows, 72, seed=10)
            e[:, 64] = 1.0
            e[:, 65:] = 0.0
            c = torch.zeros(1024, 72, device=DEV)
            native.gemm([dict(a=native.split(gfull), a_mn=True, b=native.split(e), b_mn=True, c=c, split_k=0)])
            torch.cuda.synchronize()
            report("G^T [e|1] rows %d -> 1024x72" % rows, c, gfull.double().t() @ e.double())
    if which in ("all", "mnk"):
        for rows, m, n in ((64, 128, 64), (1000, 384, 128)):
            dy, x = rand(rows, m, seed=7), rand(n, rows, seed=8)
            c = torch.zeros(m, n, device=DEV)
            native.gemm([dict(a=native.split(dy), a_mn=True, b=native.split(x), c=c)])
            torch.cuda.synchronize()
            report("MN-major x K-major rows %d -> %dx%d" % (rows, m, n), c, dy.double().t() @ x.double().t())


if __name__ == "__main__":
    main()

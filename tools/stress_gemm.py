"""Development aid: repeat small cn_gemm_bf16x3 launches back to back and report any run whose result deviates (a race would
show as a rare large error)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crowdnav_dsrnn_b200 import native  # noqa: E402

DEV = torch.device("cuda:0")


def rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    cases = []
    for rows, m, n in ((200, 64, 512), (210, 1024, 72), (42, 1024, 72), (200, 256, 256), (200, 384, 128), (1000, 768, 256)):
        dy, x = rand(rows, m, seed=1), rand(rows, n, seed=2)
        want = dy.double().t() @ x.double()
        cases.append(("wgrad rows %d -> %dx%d" % (rows, m, n), native.split(dy), native.split(x), want, dict(a_mn=True, b_mn=True, split_k=0), (m, n)))
    for m, n, k in ((200, 64, 512), (200, 256, 256), (200, 384, 128), (210, 256, 768)):
        x, w = rand(m, k, seed=3), rand(n, k, seed=4)
        cases.append(("linear %dx%dx%d" % (m, n, k), native.split(x), native.split(w), x.double() @ w.double().t(), dict(), (m, n)))
        w2 = rand(k, n, seed=5)
        cases.append(("dx %dx%dx%d" % (m, n, k), native.split(x), native.split(w2), x.double() @ w2.double(), dict(b_mn=True), (m, n)))
    bad = 0
    for it in range(iters):
        outs = []
        for name, a, b, want, kw, shape in cases:          # all launches of an iteration are enqueued before any check
            c = torch.zeros(*shape, device=DEV)
            native.gemm([dict(a=a, b=b, c=c, **kw)])
            outs.append(c)
        for (name, a, b, want, kw, shape), c in zip(cases, outs):
            err = (c.double() - want).abs().max().item()
            if err > 1e-3 * max(1.0, want.abs().max().item()):
                bad += 1
                wrong = ((c.double() - want).abs() > 1e-3 * want.abs().max()).nonzero()
                print("iter %d %s: err %.3e, %d wrong elements, rows %s cols %s" % (it, name, err, wrong.shape[0],
                      sorted(set((wrong[:, 0] // 32).tolist()))[:12], sorted(set((wrong[:, 1] // 32).tolist()))[:12]), flush=True)
    print("stress: %d bad results in %d iterations x %d cases" % (bad, iters, len(cases)))


if __name__ == "__main__":
    main()

"""One GEMM shape, a few launches: the target of an ncu capture (tools/ncu_gemm_case.py linear|recurrent)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crowdnav_dsrnn_b200 import native  # noqa: E402

DEV = torch.device("cuda:0")
pair = lambda *s: (torch.randn(*s, device=DEV).bfloat16(), (torch.randn(*s, device=DEV) * 1e-3).bfloat16())
which = sys.argv[1] if len(sys.argv) > 1 else "linear"
if which == "linear":
    x, w, y = pair(122880, 256), pair(256, 256), torch.empty(122880, 256, device=DEV)
    fn = lambda: native.gemm([dict(a=x, b=w, c=y)])
else:
    g, w, d = pair(86016, 1024), pair(768, 256), torch.zeros(86016, 256, device=DEV)
    fn = lambda: native.gemm([dict(a=(g[0][:, 256:], g[1][:, 256:]), b=w, b_mn=True, c=d, accumulate=True)])
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("done", which)

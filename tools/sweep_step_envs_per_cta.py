"""Crowd-step kernel at the c3 shape for several envs-per-CTA (CN_STEP_ENVS_PER_CTA is read at every launch): the grid is
ceil(N / E) CTAs on 148 SMs x 4 resident CTAs, so E decides how full the last wave is."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crowdnav_dsrnn_b200 import _lib  # noqa: E402
from crowdnav_dsrnn_b200.engine import CrowdEngine  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda:0")
    wl = bench.WORKLOADS["c3"]
    N = int(sys.argv[1]) if len(sys.argv) > 1 else wl["envs_per_gpu"]
    cfg = bench.make_config(wl)
    eng = CrowdEngine(cfg, N, dev, phase="train")
    eng.reset()
    g = torch.Generator(device=dev).manual_seed(0)
    act = torch.randn(N, 2, device=dev, generator=g) * 0.5
    for _ in range(150):
        eng.step(act, auto_reset=True)
    for split in ("0", "1", "0", "1"):
        os.environ["CN_STEP_SPLIT"] = split
        os.environ.pop("CN_STEP_ENVS_PER_CTA", None)
        for _ in range(10):
            eng.step(act, auto_reset=True)
        lib.cn_env_enable_timing(eng.handle, 1)
        for _ in range(40):
            eng.step(act, auto_reset=True)
        ms, n = C.c_float(), C.c_int()
        lib.cn_env_time_ms(eng.handle, C.byref(ms), C.byref(n))
        lib.cn_env_enable_timing(eng.handle, 0)
        print("CN_STEP_SPLIT=%s  step %.4f ms" % (split, ms.value / n.value), flush=True)
    os.environ.pop("CN_STEP_SPLIT", None)
    for E in (0, 6, 7, 8, 10):
        if E:
            os.environ["CN_STEP_ENVS_PER_CTA"] = str(E)
        else:
            os.environ.pop("CN_STEP_ENVS_PER_CTA", None)
        for _ in range(10):
            eng.step(act, auto_reset=True)
        lib.cn_env_enable_timing(eng.handle, 1)
        for _ in range(40):
            eng.step(act, auto_reset=True)
        ms, n = C.c_float(), C.c_int()
        lib.cn_env_time_ms(eng.handle, C.byref(ms), C.byref(n))
        lib.cn_env_enable_timing(eng.handle, 0)
        t = ms.value / n.value
        e = E or 8
        ctas = (N + e - 1) // e
        print("E=%2s  %5d CTAs = %.2f waves of 592   step kernel %.4f ms" % (E or "def", ctas, ctas / 592.0, t), flush=True)


if __name__ == "__main__":
    main()

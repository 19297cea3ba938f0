"""Crowd-step kernel (K1) alone at several (H, N): CUDA-event time per launch, algorithmic GB/s (B_step(H) = 96 H + 188)
and fraction of the measured HBM peak.  State is made larger than L2 so HBM traffic is real."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crowdnav_dsrnn_b200 import Config, _lib  # noqa: E402
from crowdnav_dsrnn_b200.engine import CrowdEngine  # noqa: E402


def main():
    lib = _lib.load()
    peak = bench.measured_peaks()["hbm_gbs"]
    dev = torch.device("cuda:0")
    for H, N, kin in ((1, 1 << 21, "holonomic"), (3, 1 << 20, "holonomic"), (5, 1 << 20, "holonomic"), (5, 1 << 18, "unicycle"),
                      (10, 1 << 19, "holonomic"), (20, 1 << 18, "holonomic"), (20, 16384, "holonomic")):
        cfg = Config(kinematics=kin, human_num=H)
        eng = CrowdEngine(cfg, N, dev, phase="train")
        eng.reset()
        act = torch.randn(N, 2, device=dev) * (0.05 if kin == "unicycle" else 0.5)
        for _ in range(30):
            eng.step(act, auto_reset=True)
        lib.cn_env_enable_timing(eng.handle, 1)
        for _ in range(20):
            eng.step(act, auto_reset=True)
        ms, n = C.c_float(), C.c_int()
        lib.cn_env_time_ms(eng.handle, C.byref(ms), C.byref(n))
        t = ms.value / n.value
        gbs = N * bench.step_bytes(H) / (t * 1e-3) / 1e9
        print("H=%2d N=%8d %-9s step kernel %8.3f ms  %7.1f M env-steps/s  %7.1f GB/s algorithmic = %.3f of %.0f GB/s"
              % (H, N, kin, t, N / t / 1e3, gbs, gbs / peak, peak), flush=True)
        eng.close()
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

"""Per-role cycle breakdown of edge_gru_tc_kernel. Needs a profiling build:
    make -C crowdnav_dsrnn_b200/csrc clean all NET_FLAGS=-DEDGE_PROFILE
Development aid; rebuild without the flag afterwards."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crowdnav_dsrnn_b200 import _lib  # noqa: E402
from crowdnav_dsrnn_b200.model import Policy  # noqa: E402
from crowdnav_dsrnn_b200.spaces import crowd_spaces  # noqa: E402

NAMES = {0: "stage", 6: "stage: wait a_free", 1: "epi: wait tmem_full", 2: "epi: ld+gates+store", 3: "epi: hprev+bar", 4: "epi warps total", 5: "ctas",
         8: "mma: wait a_ready", 9: "mma: wait tmem_empty", 10: "mma: wait B full", 11: "mma: chunk commits", 12: "mma total",
         16: "producer: wait B empty", 17: "producer total"}


def main():
    N, H = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 20
    iters = 10
    dev = torch.device("cuda:0")
    lib = _lib.load()
    fn = lib.cn_debug_edge_profile
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    wl = bench.WORKLOADS["c3"]
    cfg = bench.make_config(wl)
    obs_space, act_space = crowd_spaces(H)
    policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg)
    policy.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(wl["weights"]).items()})
    policy = policy.to(dev)
    obs = {"robot_node": torch.randn(N, 1, 7, device=dev), "temporal_edges": torch.randn(N, 1, 2, device=dev),
           "spatial_edges": torch.randn(N, H, 2, device=dev)}
    hx = {"human_node_rnn": torch.randn(N, 1, 128, device=dev) * 0.3, "human_human_edge_rnn": torch.randn(N, H + 1, 256, device=dev) * 0.3}
    masks = torch.ones(N, 1, device=dev)
    # resident-image path (timing only: the input image holds arbitrary values): argv[3] = none | in | out | both
    img_mode = sys.argv[3] if len(sys.argv) > 3 else "none"
    imgs = [tuple((torch.randn(N * (H + 1), 256, device=dev) * s).to(torch.bfloat16) for s in (0.3, 0.001)) for _ in range(2)]
    for prec in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["bf16x3", "fp16"]):
        policy.precision = prec
        for _ in range(3):
            policy.act(obs, dict(hx), masks, deterministic=True)
        if img_mode != "none" and prec == "bf16x3":
            policy.set_edge_image(imgs[0] if img_mode in ("in", "both") else None, imgs[1] if img_mode in ("out", "both") else None)
            print("== resident image: %s" % img_mode)
        torch.cuda.synchronize()
        fn(None, 1)
        for _ in range(iters):
            policy.act(obs, dict(hx), masks, deterministic=True)
        torch.cuda.synchronize()
        out = (C.c_ulonglong * 32)()
        fn(out, 1)
        ctas = out[5] / iters
        tiles = ((N * H + 127) // 128 + (N + 127) // 128)
        print("== %s: %d CTAs, %.1f tiles per CTA; per-CTA per-launch microseconds at 1.965 GHz (per tile in brackets)" % (prec, ctas, tiles / ctas))
        for i, name in NAMES.items():
            if i == 5:
                continue
            us = out[i] / iters / (ctas / 2 if 8 <= i <= 12 else ctas) / 1965.0      # MMA counters: leader CTAs only
            print("  %-26s %9.1f us  [%6.2f]" % (name, us, us / (tiles / ctas)))


if __name__ == "__main__":
    main()

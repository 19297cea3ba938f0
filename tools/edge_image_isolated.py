"""Forward-only ping-pong loop (no crowd step) with the resident edge image, CUPTI kernel durations of the edge kernel.
usage: python tools/edge_image_isolated.py [image: 1|0] [zero-mask fraction] [n_envs] [chain: 1|0]
chain=0: every forward reads the SAME random-normal hidden state (and its split-bf16 image) instead of the recurrent chain.
Separates what the edge kernel costs by itself from what it costs inside the rollout (profiles/r2_edge_resident_image.txt)."""
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from crowdnav_dsrnn_b200.model import Policy  # noqa: E402
from crowdnav_dsrnn_b200.spaces import crowd_spaces  # noqa: E402


def main():
    use_img = (sys.argv[1] if len(sys.argv) > 1 else "1") == "1"
    zero_frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
    chain = (sys.argv[4] if len(sys.argv) > 4 else "1") == "1"
    wl = bench.WORKLOADS["c3"]
    H = wl["human_num"]
    dev = torch.device("cuda:0")
    cfg = bench.make_config(wl)
    obs_space, act_space = crowd_spaces(H)
    policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg)
    policy.load_state_dict({k: torch.from_numpy(v) for k, v in bench.load_weights(wl["weights"]).items()})
    policy = policy.to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    obs = {"robot_node": torch.randn(n, 1, 7, device=dev, generator=g), "temporal_edges": torch.randn(n, 1, 2, device=dev, generator=g),
           "spatial_edges": torch.randn(n, H, 2, device=dev, generator=g) * 3}
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
    sets = [dict(h_node=z(n, 1, 128), h_edge=z(n, H + 1, 256), value=z(n, 1), mean=z(n, 2),
                 image=tuple(torch.zeros(n * (H + 1), 256, dtype=torch.bfloat16, device=dev) for _ in range(2))) for _ in range(2)]
    masks = [(torch.rand(n, 1, device=dev, generator=g) >= zero_frac).float() for _ in range(2)]

    if not chain:       # a fixed random-normal state with its exact split image in set 1; every forward reads it, writes set 0
        h = torch.randn(n, H + 1, 256, device=dev, generator=g) * 0.3
        sets[1]["h_edge"].copy_(h)
        sets[1]["h_node"].copy_(torch.randn(n, 1, 128, device=dev, generator=g) * 0.3)
        logical = torch.cat([h[:, 1:].reshape(n * H, 256), h[:, 0]], 0)
        hi = logical.to(torch.bfloat16)
        sets[1]["image"][0].copy_(hi)
        sets[1]["image"][1].copy_((logical - hi.float()).to(torch.bfloat16))

    def fwd(p, first=False):
        if not chain:
            p, first = 0, False
        a, b = sets[p], sets[p ^ 1]
        if use_img:
            policy.set_edge_image(None if first else b["image"], a["image"])
        policy.cuda_forward(obs, {"human_node_rnn": b["h_node"], "human_human_edge_rnn": b["h_edge"]}, masks[p], need_features=False,
                            out=dict(h_node=a["h_node"], h_edge=a["h_edge"], value=b["value"], mean=b["mean"]))

    with torch.no_grad():
        policy.cuda_forward(obs, {"human_node_rnn": sets[1]["h_node"], "human_human_edge_rnn": sets[1]["h_edge"]}, masks[0], need_features=False)
        if chain:
            fwd(0, first=True)
        p = 1
        for _ in range(150):
            fwd(p)
            p ^= 1
        torch.cuda.synchronize()
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            for _ in range(6):
                fwd(p)
                p ^= 1
            torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel" and "edge_gru" in e["name"]]
    print("image=%d zero_frac=%.3f n=%d chain=%d: edge kernel us:" % (use_img, zero_frac, n, chain), ["%.1f" % e["dur"] for e in ev], ev[0]["name"][:70])


if __name__ == "__main__":
    main()

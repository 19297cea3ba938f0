"""Times the training-mode edge kernel alone (cn_dsrnn_edge_sequence_step, T launches) -- development aid.
    CROWDNAV_B200_LIB=<EDGE_PROFILE build> CN_EDGE_DEBUG=16|32|48 python tools/bench_edge_train.py   # what-if: no ws / record stores"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crowdnav_dsrnn_b200 import Config, native  # noqa: E402
from crowdnav_dsrnn_b200.model import Policy  # noqa: E402
from crowdnav_dsrnn_b200.spaces import crowd_spaces  # noqa: E402

n, H, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 20, 30
dev = torch.device("cuda:0")
cfg = Config(human_num=H)
obs_space, act_space = crowd_spaces(H)
policy = Policy(obs_space.spaces, act_space, base="srnn", base_kwargs=cfg).to(dev)
se, te = torch.randn(T, n * H, 2, device=dev), torch.randn(T, n, 2, device=dev)
h0 = torch.randn(n * H + n, 256, device=dev) * 0.3
masks = (torch.rand(T, n, device=dev) > 0.03).float()
es, et = policy.base.humanhumanEdgeRNN_spatial, policy.base.humanhumanEdgeRNN_temporal
params = [p for m in (es, et) for p in (m.encoder_linear.weight, m.encoder_linear.bias, m.gru.weight_ih_l0, m.gru.weight_hh_l0,
                                        m.gru.bias_ih_l0, m.gru.bias_hh_l0)]
with torch.no_grad():
    for _ in range(2):
        native.EdgeGruSequence.apply(policy, se, te, h0, masks, *params)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        native.EdgeGruSequence.apply(policy, se, te, h0, masks, *params)
    e1.record()
    torch.cuda.synchronize()
print("edge sequence forward, %d envs x %d humans x %d steps: %.3f ms per sequence (%.1f us per step) [CN_EDGE_DEBUG=%s]"
      % (n, H, T, e0.elapsed_time(e1) / 3, e0.elapsed_time(e1) / 3 / T * 1e3, os.environ.get("CN_EDGE_DEBUG", "0")))

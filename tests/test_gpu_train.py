"""GPU: the training loop -- CUDA rollout + torch PPO update on the device, reference artefact
names, and the CUDA forward picking up the weights Adam just changed."""
import csv
import math
import os

import numpy as np
import pytest
import torch

from crowdnav_dsrnn_b200 import Config
from crowdnav_dsrnn_b200.model import Policy
from crowdnav_dsrnn_b200.spaces import crowd_spaces
from crowdnav_dsrnn_b200.train import train

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _shipped_policy(cfg):
    w = np.load(os.path.join(GOLDEN, "weights_holonomic_27776.npz"))
    obs, act = crowd_spaces(cfg.sim.human_num)
    p = Policy(obs.spaces, act, base="srnn", base_kwargs=cfg)
    p.load_state_dict({k: torch.from_numpy(w[k]) for k in w.files})
    return p.to(DEV)


def test_training_loop_runs_and_writes_reference_artefacts(tmp_path):
    cfg = Config()
    cfg.training.num_processes = 256
    cfg.training.log_interval = 1
    cfg.training.save_interval = 2
    policy = _shipped_policy(cfg)
    before = {k: v.detach().clone() for k, v in policy.state_dict().items()}
    policy2, history = train(cfg, DEV, num_updates=4, output_dir=str(tmp_path), actor_critic=policy, log=None,
                             max_envs_per_pass=64)
    assert policy2 is policy and len(history) == 4
    for row in history:
        assert all(math.isfinite(row[k]) for k in ("loss/value_loss", "loss/policy_loss", "loss/policy_entropy"))
    # the shipped policy keeps solving the task while being fine-tuned with sampled actions (reference: 0.93 deterministic)
    episodes = sum(r["episodes"] for r in history)
    success = sum(r["success"] * r["episodes"] for r in history) / episodes
    assert episodes > 300 and success > 0.75, (episodes, success)
    changed = [k for k, v in policy.state_dict().items() if not torch.equal(v, before[k])]
    unused = [k for k in before if k not in changed]      # state_dict entries the DS-RNN forward never reads get no gradient
    assert len(changed) >= 40 and all(dict(policy.named_parameters())[k].grad is None for k in unused), unused
    assert sorted(os.listdir(tmp_path / "checkpoints")) == ["00000.pt", "00002.pt", "00003.pt"]
    sd = torch.load(tmp_path / "checkpoints" / "00003.pt", map_location="cpu")
    assert sorted(sd) == sorted(before) and torch.equal(sd["dist.logstd._bias"], policy.state_dict()["dist.logstd._bias"].cpu())
    with open(tmp_path / "progress.csv") as f:
        rows = list(csv.DictReader(f))
    assert list(rows[0]) == ["misc/nupdates", "misc/total_timesteps", "fps", "eprewmean", "loss/policy_entropy",
                             "loss/policy_loss", "loss/value_loss"] and len(rows) == 4
    assert int(rows[-1]["misc/total_timesteps"]) == 4 * 256 * 30
    # train.py:47-75: config snapshot (what test.py re-imports as <model_dir>.configs.train_config) and output.log
    assert os.path.exists(tmp_path / "configs" / "train_config.py") and os.path.lexists(tmp_path / "configs" / "config.py")
    assert "training.num_processes = 256" in (tmp_path / "configs" / "train_config_values.txt").read_text()
    assert (tmp_path / "output.log").read_text().count("Updates ") == 4


def test_test_cli_writes_the_reference_log(tmp_path, monkeypatch):
    """`python -m crowdnav_dsrnn_b200.test` (test.py:25-214): flags, model-directory layout, log file name, log lines."""
    import re

    from crowdnav_dsrnn_b200 import test as test_cli

    w = np.load(os.path.join(GOLDEN, "weights_holonomic_27776.npz"))
    ckpt_dir = tmp_path / "data" / "m" / "checkpoints"
    ckpt_dir.mkdir(parents=True)
    torch.save({k: torch.from_numpy(w[k]) for k in w.files}, ckpt_dir / "27776.pt")
    monkeypatch.chdir(tmp_path)
    assert test_cli.main(["--model_dir", "data/m", "--test_model", "27776.pt", "--test_name", "cli"]) == 0
    log = (tmp_path / "data" / "m" / "test" / "model_27776_test_cli_.log").read_text()
    assert "Using model" in log and "robot FOV" in log
    rates = [float(re.search(r"%s rate: ([0-9.]+)" % k, log).group(1)) for k in ("success", "collision", "timeout")]
    assert abs(sum(rates) - 1.0) < 2e-3 and rates[0] > 0.85          # reference: 0.93 / 0.055 / 0.012 over 2000 episodes
    with pytest.raises(NotImplementedError):
        test_cli.main(["--model_dir", "data/m", "--viz"])


def test_cuda_forward_tracks_optimizer_updates():
    """Adam updates the parameters in place; the next `act` must run on the new weights (re-packed tcgen05 images)."""
    cfg = Config()
    policy = _shipped_policy(cfg)
    n, H = 32, 5
    g = torch.Generator(device="cpu").manual_seed(3)
    obs = {"robot_node": torch.randn(n, 1, 7, generator=g).to(DEV), "temporal_edges": torch.randn(n, 1, 2, generator=g).to(DEV),
           "spatial_edges": torch.randn(n, H, 2, generator=g).to(DEV)}
    hx = lambda: {"human_node_rnn": torch.zeros(n, 1, 128, device=DEV), "human_human_edge_rnn": torch.zeros(n, H + 1, 256, device=DEV)}
    masks = torch.ones(n, 1, device=DEV)
    v0, a0, _, _ = policy.act(obs, hx(), masks, deterministic=True)
    opt = torch.optim.Adam(policy.parameters(), lr=1e-2)
    value, _, _, _ = policy.evaluate_actions(obs, hx(), masks, a0)
    value.mean().backward()
    opt.step()
    v1, a1, _, _ = policy.act(obs, hx(), masks, deterministic=True)
    vt, _, _, _ = policy.evaluate_actions(obs, hx(), masks, a1)
    assert (v1 - v0).abs().max() > 1e-3                    # the weights moved ...
    assert (v1 - vt.detach()).abs().max() < 1e-3 * max(1.0, float(vt.abs().max()))   # ... and CUDA and torch paths agree on them


def test_batched_sequence_forward_equals_per_step_graph_on_gpu():
    """On CUDA the masked GRU sequence uses ATen's fused GRU-cell pair for the gate math (forward and backward)."""
    from test_config_host import sequence_impls_agree

    sequence_impls_agree(torch.float32, DEV, 5e-4)


@pytest.mark.parametrize("rows,hid", [(1, 128), (77, 256), (4099, 256)])
def test_gru_gate_kernels_match_torch_reference(rows, hid):
    """cn_gru_gates_forward / _backward through the C ABI against the same step written with torch fp32 ops + autograd."""
    import ctypes as C

    from crowdnav_dsrnn_b200 import _lib

    lib = _lib.load()
    ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    g = torch.Generator(device="cpu").manual_seed(rows + hid)
    mk = lambda *shape: torch.randn(*shape, generator=g).to(DEV)
    gi, gh, h_prev = mk(rows, 3 * hid).requires_grad_(True), mk(rows, 3 * hid).requires_grad_(True), mk(rows, hid)
    b_ih, b_hh = mk(3 * hid), mk(3 * hid)
    m_t = (torch.rand(rows, 1, generator=g) > 0.3).float().to(DEV)
    m_next = (torch.rand(rows, 1, generator=g) > 0.3).float().to(DEV)
    hm = (h_prev * m_t).requires_grad_(True)
    # torch reference of the step (gate order r|z|n, torch.nn.GRU)
    a, b = gi + b_ih, gh + b_hh
    r = torch.sigmoid(a[:, :hid] + b[:, :hid])
    z = torch.sigmoid(a[:, hid:2 * hid] + b[:, hid:2 * hid])
    n = torch.tanh(a[:, 2 * hid:] + r * b[:, 2 * hid:])
    h_ref = n + z * (hm - n)
    grad_h, d_next = mk(rows, hid), mk(rows, hid)
    h_ref.backward(grad_h + d_next * m_next)
    stream = C.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)
    h_out, hm_next, ws = torch.empty(rows, hid, device=DEV), torch.empty(rows, hid, device=DEV), torch.empty(rows, 4 * hid, device=DEV)
    _lib.check(lib.cn_gru_gates_forward(ptr(gi.detach()), ptr(gh.detach()), ptr(hm.detach()), ptr(b_ih), ptr(b_hh), ptr(m_next),
                                        ptr(h_out), ptr(hm_next), ptr(ws), None, None, rows, hid, stream), "cn_gru_gates_forward")
    dgi, dgh, dhm = torch.empty(rows, 3 * hid, device=DEV), torch.empty(rows, 3 * hid, device=DEV), torch.empty(rows, hid, device=DEV)
    _lib.check(lib.cn_gru_gates_backward(ptr(grad_h), ptr(d_next), ptr(m_next), ptr(ws), ptr(hm.detach()), ptr(dgi), ptr(dgh),
                                         ptr(dhm), None, None, None, None, rows, hid, stream), "cn_gru_gates_backward")
    torch.cuda.synchronize()
    tol = 2e-5      # fp32 with fused multiply-adds and libdevice expf / tanhf vs ATen's
    for got, want in ((h_out, h_ref.detach()), (hm_next, h_ref.detach() * m_next), (ws[:, :hid], r.detach()),
                      (ws[:, hid:2 * hid], z.detach()), (ws[:, 2 * hid:3 * hid], n.detach()), (dgi, gi.grad), (dgh, gh.grad),
                      (dhm, hm.grad)):
        assert (got - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())
    # bad arguments are refused, not launched
    assert lib.cn_gru_gates_forward(ptr(gi.detach()), ptr(gh.detach()), ptr(hm.detach()), ptr(b_ih), ptr(b_hh), None, ptr(h_out),
                                    ptr(hm_next), ptr(ws), None, None, rows, hid, stream) == -1
    assert lib.cn_gru_gates_backward(ptr(grad_h), ptr(d_next), ptr(m_next), ptr(ws), ptr(hm.detach()), ptr(dgi), ptr(dgh), ptr(dhm),
                                     None, None, None, None, rows, 6, stream) == -1
    # the optional bf16 pairs are the split of the fp32 outputs
    bf = lambda *shape: torch.empty(*shape, dtype=torch.bfloat16, device=DEV)
    nh, nl = bf(rows, hid), bf(rows, hid)
    _lib.check(lib.cn_gru_gates_forward(ptr(gi.detach()), ptr(gh.detach()), ptr(hm.detach()), ptr(b_ih), ptr(b_hh), ptr(m_next),
                                        ptr(h_out), ptr(hm_next), ptr(ws), ptr(nh), ptr(nl), rows, hid, stream), "cn_gru_gates_forward")
    pairs = [bf(rows, 3 * hid) for _ in range(4)]
    _lib.check(lib.cn_gru_gates_backward(ptr(grad_h), ptr(d_next), ptr(m_next), ptr(ws), ptr(hm.detach()), ptr(dgi), ptr(dgh),
                                         ptr(dhm), *[ptr(t) for t in pairs], rows, hid, stream), "cn_gru_gates_backward")
    torch.cuda.synchronize()
    for full, hi, lo in ((hm_next, nh, nl), (dgi, pairs[0], pairs[1]), (dgh, pairs[2], pairs[3])):
        assert torch.equal(hi, full.to(torch.bfloat16))
        assert ((hi.float() + lo.float() - full).abs() <= full.abs() * 2.0 ** -16 + 1e-38).all()


def test_bf16x3_recurrent_gemms_keep_fp32_level_accuracy(monkeypatch):
    """PPO(bf16x3=True): the recurrent products of the masked GRU sequences as split-bf16 3-pass tensor-core GEMMs."""
    from crowdnav_dsrnn_b200 import model as model_mod
    from test_config_host import sequence_impls_agree

    a = torch.randn(1000, 260, device=DEV) * torch.logspace(-3, 3, 260, device=DEV)
    hi, lo = model_mod._split_bf16(a)                               # cn_split_bf16
    assert hi.dtype is torch.bfloat16 and torch.equal(hi, a.to(torch.bfloat16))
    assert ((hi.float() + lo.float() - a).abs() <= a.abs() * 2.0 ** -16).all()
    monkeypatch.setattr(model_mod, "SEQUENCE_GEMM", "bf16x3")      # only the batched form reads it; per_step stays fp32
    sequence_impls_agree(torch.float32, DEV, 1e-3)

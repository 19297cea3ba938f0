"""GPU: the native PPO-update path -- cn_gemm_bf16x3 (TMA-fed tcgen05 GEMM, K-major / MN-major operands,
grouped problems, split-K), the edge-GRU sequence kernels, and `Policy.evaluate_actions(sequence_impl="native")` +
`PPO(native=True)` against the torch restatement and against the reference's own PPO.update golden."""
import os

import numpy as np
import pytest
import torch

from crowdnav_dsrnn_b200 import native
from crowdnav_dsrnn_b200.ppo import PPO

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _check(got, want, tol=2e-5):
    err = (got.double() - want).abs().max().item()
    assert err <= tol * max(1.0, want.abs().max().item()), (err, want.abs().max().item())


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (300, 80, 200), (1, 16, 8), (1000, 320, 768), (4099, 384, 128), (257, 512, 256)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_kmajor_forward_linear(m, n, k, act):
    """C = act(A B^T + bias), A [m, k], B [n, k] both K-major (the forward of a linear layer)."""
    x, w, b = _rand(m, k, seed=1), _rand(n, k, seed=2, scale=k ** -0.5), _rand(n, seed=3)
    y = torch.full((m, n), float("nan"), device=DEV)
    native.gemm([dict(a=native.split(x), b=native.split(w), c=y, bias=b, act=act)])
    want = x.double() @ w.double().t() + b.double()
    want = torch.relu(want) if act == 1 else torch.tanh(want) if act == 2 else want
    _check(y, want, tol=5e-5)


@pytest.mark.parametrize("m,n,k", [(256, 256, 768), (77, 64, 768), (513, 128, 384), (130, 200, 72)])
def test_gemm_mn_major_b_and_accumulate(m, n, k):
    """C += A B with B [k, n] MN-major (dx = dy W) on a strided A (a column slice of a wider array)."""
    wide = _rand(m, k + 256, seed=4)
    a = wide[:, 256:]
    w = _rand(k, n, seed=5, scale=k ** -0.5)
    hi, lo = native.split(wide)
    c0 = _rand(m, n, seed=6)
    c = c0.clone()
    native.gemm([dict(a=(hi[:, 256:], lo[:, 256:]), b=native.split(w), b_mn=True, c=c, accumulate=True)])
    _check(c, c0.double() + a.double() @ w.double())
    c2 = torch.empty(m, n, device=DEV)
    native.gemm([dict(a=(hi[:, 256:], lo[:, 256:]), b=native.split(w), b_mn=True, c=c2)])
    _check(c2, a.double() @ w.double())


@pytest.mark.parametrize("rows,m,n,split", [(5000, 768, 256, 0), (5000, 768, 64, 0), (130, 384, 128, 1), (70000, 256, 512, 0), (999, 100 * 8, 72, 3)])
def test_gemm_weight_gradient_both_mn_major_split_k(rows, m, n, split):
    """dW = dY^T X: A = dY [rows, m] and B = X [rows, n] both MN-major, reduction over the rows cut into k-slices."""
    dy, x = _rand(rows, m, seed=7), _rand(rows, n, seed=8)
    c = torch.zeros(m, n, device=DEV)
    native.gemm([dict(a=native.split(dy), a_mn=True, b=native.split(x), b_mn=True, c=c, split_k=split)])
    want = dy.double().t() @ x.double()
    _check(c, want, tol=3e-5)


def test_gemm_grouped_problems_and_single_pass_operands():
    """Four problems of different shape / major in one launch; lo = None drops that operand's correction pass."""
    xs = [_rand(500, 128, seed=10), _rand(64, 768, seed=11), _rand(3000, 384, seed=12), _rand(300, 64, seed=13)]
    ws = [_rand(256, 128, seed=14), _rand(768, 256, seed=15), _rand(3000, 128, seed=16), _rand(320, 64, seed=17)]
    cs = [torch.empty(500, 256, device=DEV), torch.empty(64, 256, device=DEV), torch.zeros(384, 128, device=DEV), torch.empty(300, 320, device=DEV)]
    native.gemm([dict(a=native.split(xs[0]), b=native.split(ws[0]), c=cs[0]),
                 dict(a=native.split(xs[1]), b=native.split(ws[1]), b_mn=True, c=cs[1]),
                 dict(a=native.split(xs[2]), a_mn=True, b=native.split(ws[2]), b_mn=True, c=cs[2], split_k=0),
                 dict(a=native.split(xs[3]), b=native.split(ws[3]), c=cs[3])])
    _check(cs[0], xs[0].double() @ ws[0].double().t())
    _check(cs[1], xs[1].double() @ ws[1].double())
    _check(cs[2], xs[2].double().t() @ ws[2].double(), tol=3e-5)
    _check(cs[3], xs[3].double() @ ws[3].double().t())
    x, w = xs[0], ws[0]
    c = torch.empty(500, 256, device=DEV)
    native.gemm([dict(a=(x.bfloat16(), None), b=(w.bfloat16(), None), c=c)])
    _check(c, x.bfloat16().double() @ w.bfloat16().double().t(), tol=1e-5)


def test_gemm_refuses_bad_arguments():
    x, w = native.split(_rand(64, 256)), native.split(_rand(64, 256))
    c = torch.empty(64, 64, device=DEV)
    with pytest.raises(Exception):
        native.gemm([dict(a=x, b=w, c=c, bias=_rand(64), split_k=2)])        # split-K cannot carry a bias
    with pytest.raises(Exception):
        native.gemm([dict(a=(x[0][:, :60], x[1][:, :60]), b=w, c=c)])        # k mismatch
    with pytest.raises(Exception):
        native.gemm([dict(a=x, b=w, c=c)] * 5)                               # too many problems


@pytest.mark.parametrize("act", [None, "relu", "tanh"])
def test_native_linear_autograd_matches_torch(act):
    lin = torch.nn.Linear(256, 64).to(DEV)
    x = _rand(700, 256, seed=20).requires_grad_(True)
    y = native.linear(x, lin, act)
    gy = _rand(700, 64, seed=21)
    if act == "relu":            # no gradient through outputs so close to the kink that fp32 and fp64 may disagree on the side
        gy = gy * (torch.nn.functional.linear(x, lin.weight, lin.bias).abs() > 1e-3)
    y.backward(gy)
    got = (y.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = None
    lin.zero_grad()
    f = {None: lambda t: t, "relu": torch.relu, "tanh": torch.tanh}[act]
    y2 = f(torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double()))
    gx, gw, gb = torch.autograd.grad(y2, (x, lin.weight, lin.bias), gy.double())
    for a, b in zip(got, (y2.detach(), gx, gw, gb)):
        _check(a, b.double(), tol=5e-5)


@pytest.mark.parametrize("case", [dict(n=6, H=5, T=7), dict(n=40, H=20, T=5, seed=3), dict(n=130, H=3, T=4, seed=5)])
def test_native_sequence_forward_equals_per_step_graph(case):
    """evaluate_actions on the native kernels (edge GRU sequence kernel, native GEMMs everywhere) against the per-step torch
    fp32 autograd graph: values, log-probs, final hidden states, every parameter gradient, gradient of the initial state."""
    from test_config_host import sequence_impls_agree

    sequence_impls_agree(torch.float32, DEV, 1e-3, impls=("per_step", "native"), case=case)


def test_ppo_update_native_matches_reference_golden_on_cuda():
    """The reference's OWN PPO.update (oracle/gen_golden_ppo.py) vs PPO(native=True) on the GPU: losses and every
    parameter's change after 5 epochs x 2 minibatches within the same 2 % as the CPU path."""
    from test_ppo_update import _check_update, _golden, _policy, _storage

    g = _golden()
    hyper = g["hyper"]
    policy = _policy().to(DEV)
    before = {k: v.detach().cpu().clone() for k, v in policy.state_dict().items()}
    st = _storage(g, keep_hidden_history=False).to(DEV)
    st.compute_returns(torch.from_numpy(g["next_value"]).to(DEV), True, float(hyper[8]), float(hyper[9]), False)
    agent = PPO(policy, float(hyper[0]), int(hyper[1]), int(hyper[2]), float(hyper[3]), float(hyper[4]), lr=float(hyper[5]),
                eps=float(hyper[6]), max_grad_norm=float(hyper[7]), native=True)
    torch.manual_seed(int(hyper[10]))
    launches = native.COUNTERS["gemm_launches"]
    losses = agent.update(st)
    assert native.COUNTERS["gemm_launches"] > launches + 100           # the native kernels really ran
    cpu_policy = _policy()
    cpu_policy.load_state_dict({k: v.detach().cpu() for k, v in policy.state_dict().items()})
    _check_update(g, cpu_policy, before, losses)


@pytest.mark.parametrize("B,H", [(3, 1), (130, 5), (1000, 20), (7, 32)])
def test_attention_mix_matches_torch_autograd(B, H):
    """cn_attention_train_forward / _backward against softmax attention written with torch ops in float64."""
    o = _rand(B, H, 256, seed=30).requires_grad_(True)
    qt = _rand(B, 256, seed=31, scale=0.2).requires_grad_(True)
    cst = _rand(B, seed=32).requires_grad_(True)
    scale = H / 8.0
    c = native.AttentionMix.apply(o, qt, cst, scale)
    gc = _rand(B, 256, seed=33)
    c.backward(gc)
    got = (c.detach(), o.grad.clone(), qt.grad.clone(), cst.grad.clone())
    od, qd, cd = o.detach().double().requires_grad_(True), qt.detach().double().requires_grad_(True), cst.detach().double().requires_grad_(True)
    alpha = torch.softmax(((od * qd.unsqueeze(1)).sum(-1) + cd.unsqueeze(1)) * scale, dim=-1)
    ref = (alpha.unsqueeze(-1) * od).sum(1)
    g = torch.autograd.grad(ref, (od, qd, cd), gc.double())
    for a, b in zip(got, (ref.detach(),) + g):
        assert (a.double() - b).abs().max().item() <= 5e-5 * max(1.0, b.abs().max().item())


def test_native_gradients_are_reproducible_across_repeated_passes():
    """Race detector: the same native forward + backward eight times; every gradient must agree with the first pass up to
    the order of the split-K atomics (1e-4 of its scale)."""
    from test_config_host import _sequence_case

    policy, obs, masks, action, h0 = _sequence_case(torch.float32, DEV, n=40, H=20, T=5, seed=3)
    policy.sequence_impl = "native"
    first = None
    for rep in range(8):
        policy.zero_grad(set_to_none=True)
        value, logp, _, out = policy.evaluate_actions(obs, {k: v.clone() for k, v in h0.items()}, masks, action)
        (value.sum() + 0.3 * logp.sum()).backward()
        grads = {k: q.grad.clone() for k, q in policy.named_parameters() if q.grad is not None}
        if first is None:
            first = grads
            continue
        bad = ["%s (pass %d): %.3e of %.3e" % (k, rep, (grads[k] - first[k]).abs().max().item(), first[k].abs().max().item())
               for k in first if (grads[k] - first[k]).abs().max().item() > 1e-4 * max(first[k].abs().max().item(), 1e-3)]
        assert not bad, bad

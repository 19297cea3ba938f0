"""`Policy`: the DS-RNN actor-critic behind the reference's interface
(pytorchBaselines/a2c_ppo_acktr/model.py:17-104): same constructor, `act`,
`get_value`, `evaluate_actions`, `.base.nenv`, `.base.human_num`,
`.is_recurrent` and the same 45 `state_dict()` keys, so either side's
checkpoints load on the other.

Rollout path (`act`, `get_value`): one call into the CUDA forward
(`cn_dsrnn_forward`, csrc/dsrnn_forward.cu + dsrnn_edge_tc.cu) -- no CPU
fallback; tensors must live on a B200.  Batch size and human count are read
from the tensor shapes (the reference bakes them from config,
srnn_model.py:361-363,410-418).

Training path (`evaluate_actions`): the same math over the
T x N rollout chunk, differentiable: masked GRU sequences (`_MaskedGruSequence`:
cuBLAS GEMMs + the library's gate kernels, csrc/dsrnn_train.cu) and batched
torch ops for everything without a recurrence.
"""
import ctypes as C
import math
import weakref

import torch
import torch.nn as nn

from . import _lib, abi, native


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _orthogonal(module, gain=1.0):
    nn.init.orthogonal_(module.weight.data, gain=gain)
    nn.init.constant_(module.bias.data, 0)
    return module


# GEMM arithmetic of the masked GRU sequences' recurrent products (PPO update): "fp32" = torch's fp32 matmul (SIMT, or TF32
# when torch.backends.cuda.matmul.allow_tf32 is set), "bf16x3" = split-bf16 3-pass on the tensor cores with fp32 accumulation
# (A_hi B_hi + A_lo B_hi + A_hi B_lo, operand error ~2^-16: the rollout kernels' precision) through three cuBLAS calls,
# "native" = the same arithmetic in ONE launch of the library's tcgen05 kernel (cn_gemm_bf16x3, csrc/gemm_bf16x3.cu); CUDA only.
SEQUENCE_GEMM = "fp32"


def _split_bf16(a):
    """a (fp32, CUDA) -> (hi, lo) bf16 with a ~= hi + lo to 2^-16 relative; one pass over a (cn_split_bf16)."""
    a = a.contiguous()
    hi = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    lo = torch.empty_like(hi)
    if a.numel() % 4 == 0 and a.numel() > 0:
        stream = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
        _lib.check(_lib.load().cn_split_bf16(_ptr(a), _ptr(hi), _ptr(lo), a.numel(), stream), "cn_split_bf16")
    else:                                                    # sizes the kernel does not take (never on the DS-RNN shapes)
        hi = a.to(torch.bfloat16)
        lo = torch.sub(a, hi).to(torch.bfloat16)
    return hi, lo


def _mm3(a_hi, a_lo, b_hi, b_lo, add=None):
    """(a_hi + a_lo) [M,K] times (b_hi + b_lo) [K,N] (bf16 pairs) without the lo*lo term, fp32 accumulation and result (+ add)."""
    f32 = torch.float32
    out = torch.mm(a_hi, b_hi, out_dtype=f32) if add is None else torch.addmm(add, a_hi, b_hi, out_dtype=f32)
    torch.addmm(out, a_lo, b_hi, out_dtype=f32, out=out)         # accumulate in place: no copy of the fp32 result per pass
    torch.addmm(out, a_hi, b_lo, out_dtype=f32, out=out)
    return out


def _mm_bf16x3(a, b_hi, b_lo, add=None):
    a_hi, a_lo = _split_bf16(a)
    return _mm3(a_hi, a_lo, b_hi, b_lo, add)


class _MaskedGruSequence(torch.autograd.Function):
    """h_t = GRUCell(x_t, m_t * h_{t-1}) for t < T over R independent rows, returned as [T, R, hidden].

    The training-time form of one DS-RNN recurrent unit (srnn_model.py:53-104 applies the masks by cutting the sequence at
    every step where any env finished -- with thousands of envs that is every step).  One autograd node for the whole
    sequence instead of ~6 per step: the input projection of all T steps is one GEMM, the weight gradients are ONE
    [3h, T*R] x [T*R, k] GEMM each at the end of the backward (per-step weight-gradient GEMMs have a 768 x 320 output and
    fill 60 of 148 SMs), only `h W_hh^T`, the gate kernel and their transposes run per step: two launches per step in each
    direction.  On CUDA the gate math is the library's `cn_gru_gates_forward / _backward` (csrc/dsrnn_train.cu), which also
    apply the masks and write straight into the [T, ...] buffers; elsewhere (CPU tests) the same formulas in plain torch.
    """

    @staticmethod
    def forward(ctx, x, h0, m, w_ih, w_hh, b_ih, b_hh):
        T, R, hid = x.shape[0], x.shape[1], h0.shape[1]
        cuda = x.is_cuda and x.dtype is torch.float32
        nat = cuda and SEQUENCE_GEMM == "native" and hid % 8 == 0 and x.shape[2] % 8 == 0
        x3 = nat or (cuda and SEQUENCE_GEMM == "bf16x3" and (R * hid) % 4 == 0)
        x_pair = ()
        if nat:
            x_pair = native.split(x.reshape(T * R, -1))
            gi = x.new_empty(T, R, 3 * hid)
            native.gemm([dict(a=x_pair, b=native.split(w_ih), c=gi.view(T * R, 3 * hid))])
        elif x3:
            gi = _mm_bf16x3(x.reshape(T * R, -1), *_split_bf16(w_ih.t())).view(T, R, 3 * hid)
        else:
            gi = torch.matmul(x, w_ih.t())                   # [T, R, 3h]; the gate math adds both biases
        hs = x.new_empty(T, R, hid)
        hm = x.new_empty(T, R, hid)                          # masked previous state of every step
        ws = x.new_empty(T, R, 4 * hid)                      # r | z | n | W_hn hm + b_hn of every step
        m = m.contiguous()
        w_hh_t = w_hh.t()
        torch.mul(h0, m[0], out=hm[0])
        pairs = ()
        if cuda:
            lib = _lib.load()
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            b_ih, b_hh = b_ih.contiguous(), b_hh.contiguous()
            if nat:
                w_pair, gh_buf = native.split(w_hh), x.new_empty(R, 3 * hid)      # W_hh [3h, hid]: K-major B of gh = hm W_hh^T
            elif x3:
                wt_hi, wt_lo = _split_bf16(w_hh_t)
            if x3:                                           # bf16 hi/lo pair of every step's masked state, written by the gate kernel
                hm_hi = torch.empty(T, R, hid, dtype=torch.bfloat16, device=x.device)
                hm_lo = torch.empty_like(hm_hi)
                _lib.check(lib.cn_split_bf16(_ptr(hm[0]), _ptr(hm_hi[0]), _ptr(hm_lo[0]), R * hid, stream), "cn_split_bf16")
                pairs = (hm_hi, hm_lo)
            for t in range(T):
                if nat:
                    gh = gh_buf
                    native.gemm([dict(a=(hm_hi[t], hm_lo[t]), b=w_pair, c=gh)])
                else:
                    gh = _mm3(hm_hi[t], hm_lo[t], wt_hi, wt_lo) if x3 else torch.mm(hm[t], w_hh_t)
                last = t == T - 1
                _lib.check(lib.cn_gru_gates_forward(_ptr(gi[t]), _ptr(gh), _ptr(hm[t]), _ptr(b_ih), _ptr(b_hh),
                                                    None if last else _ptr(m[t + 1]), _ptr(hs[t]),
                                                    None if last else _ptr(hm[t + 1]), _ptr(ws[t]),
                                                    _ptr(hm_hi[t + 1]) if x3 and not last else None,
                                                    _ptr(hm_lo[t + 1]) if x3 and not last else None, R, hid, stream),
                           "cn_gru_gates_forward")
        else:                                                # CPU tests / float64 checks: the same formulas in torch
            for t in range(T):
                a, b = gi[t] + b_ih, torch.mm(hm[t], w_hh_t) + b_hh
                r = torch.sigmoid(a[:, :hid] + b[:, :hid])
                z = torch.sigmoid(a[:, hid:2 * hid] + b[:, hid:2 * hid])
                hn = b[:, 2 * hid:]
                n = torch.tanh(a[:, 2 * hid:] + r * hn)
                hs[t] = n + z * (hm[t] - n)
                ws[t] = torch.cat([r, z, n, hn], 1)
                if t + 1 < T:
                    torch.mul(hs[t], m[t + 1], out=hm[t + 1])
        ctx.save_for_backward(x, m, w_ih, w_hh, hm, ws, *pairs, *x_pair)
        ctx.cuda = cuda
        ctx.x3 = x3
        ctx.nat = nat
        return hs

    @staticmethod
    def backward(ctx, grad_hs):
        x, m, w_ih, w_hh, hm, ws, *pairs = ctx.saved_tensors
        T, R, hid = hm.shape
        dgi = x.new_empty(T, R, 3 * hid)
        dgh = x.new_empty(T, R, 3 * hid)
        d_next = None                                        # dL/d(masked state of step t+1)
        if ctx.cuda:
            lib = _lib.load()
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            grad_hs = grad_hs.contiguous()
            dhm = x.new_empty(R, hid)
            if ctx.nat:
                w_pair = native.split(w_hh)                  # [3h, hid] = [K, N]: MN-major B of d_next = dhm + dgh W_hh
                dhm_alt = x.new_empty(R, hid)
            if ctx.x3:
                w_hi, w_lo = (None, None) if ctx.nat else _split_bf16(w_hh)
                bf = lambda: torch.empty(T, R, 3 * hid, dtype=torch.bfloat16, device=x.device)
                dgi_hi, dgi_lo, dgh_hi, dgh_lo = bf(), bf(), bf(), bf()
            for t in range(T - 1, -1, -1):
                extra = [_ptr(b[t]) for b in (dgi_hi, dgi_lo, dgh_hi, dgh_lo)] if ctx.x3 else [None] * 4
                _lib.check(lib.cn_gru_gates_backward(_ptr(grad_hs[t]), _ptr(d_next), None if d_next is None else _ptr(m[t + 1]),
                                                     _ptr(ws[t]), _ptr(hm[t]), _ptr(dgi[t]), _ptr(dgh[t]), _ptr(dhm), *extra,
                                                     R, hid, stream), "cn_gru_gates_backward")
                if ctx.nat:
                    native.gemm([dict(a=(dgh_hi[t], dgh_lo[t]), b=w_pair, b_mn=True, c=dhm, accumulate=True)])
                    d_next, dhm, dhm_alt = dhm, dhm_alt, dhm      # ping-pong: the next gate kernel reads d_next and writes dhm
                else:
                    d_next = _mm3(dgh_hi[t], dgh_lo[t], w_hi, w_lo, add=dhm) if ctx.x3 else torch.addmm(dhm, dgh[t], w_hh)
        else:
            for t in range(T - 1, -1, -1):
                g = grad_hs[t] if d_next is None else grad_hs[t] + d_next * m[t + 1]
                r, z, n, hn = ws[t].split(hid, 1)
                dpre_n = g * (1.0 - z) * (1.0 - n * n)
                dpre_r = dpre_n * hn * r * (1.0 - r)
                dpre_z = g * (hm[t] - n) * z * (1.0 - z)
                dgi[t] = torch.cat([dpre_r, dpre_z, dpre_n], 1)
                dgh[t] = torch.cat([dpre_r, dpre_z, dpre_n * r], 1)
                d_next = torch.addmm(g * z, dgh[t], w_hh)
        dgi2, dgh2 = dgi.view(T * R, 3 * hid), dgh.view(T * R, 3 * hid)
        if ctx.nat:
            gi_pair = (dgi_hi.view(T * R, 3 * hid), dgi_lo.view(T * R, 3 * hid))
            gh_pair = (dgh_hi.view(T * R, 3 * hid), dgh_lo.view(T * R, 3 * hid))
            hm_pair = (pairs[0].view(T * R, hid), pairs[1].view(T * R, hid))
            x_pair = tuple(pairs[2:4])
            dx = None
            if ctx.needs_input_grad[0]:
                dx = x.new_empty(T, R, x.shape[2])
                native.gemm([dict(a=gi_pair, b=native.split(w_ih), b_mn=True, c=dx.view(T * R, -1))])
            dw_ih, dw_hh = torch.zeros_like(w_ih), torch.zeros_like(w_hh)
            native.gemm([dict(a=gi_pair, a_mn=True, b=x_pair, b_mn=True, c=dw_ih, split_k=0),
                         dict(a=gh_pair, a_mn=True, b=hm_pair, b_mn=True, c=dw_hh, split_k=0)])
        elif ctx.x3:
            a_hi, a_lo = dgi_hi.view(T * R, 3 * hid), dgi_lo.view(T * R, 3 * hid)
            dx = _mm3(a_hi, a_lo, *_split_bf16(w_ih)).view(T, R, -1) if ctx.needs_input_grad[0] else None
            dw_ih = _mm3(a_hi.t(), a_lo.t(), *_split_bf16(x.reshape(T * R, -1)))
            dw_hh = _mm3(dgh_hi.view(T * R, 3 * hid).t(), dgh_lo.view(T * R, 3 * hid).t(),
                         pairs[0].view(T * R, hid), pairs[1].view(T * R, hid))
        else:
            dx = torch.matmul(dgi, w_ih) if ctx.needs_input_grad[0] else None
            dw_ih = torch.mm(dgi2.t(), x.reshape(T * R, -1))
            dw_hh = torch.mm(dgh2.t(), hm.view(T * R, hid))
        return dx, d_next * m[0], None, dw_ih, dw_hh, dgi2.sum(0), dgh2.sum(0)


class _EdgeRNN(nn.Module):
    """Parameter holder of HumanHumanEdgeRNN (srnn_model.py:176-215): Linear(2,64) + GRU(64,256)."""

    def __init__(self, cfg):
        super().__init__()
        self.gru = nn.GRU(cfg.human_human_edge_embedding_size, cfg.human_human_edge_rnn_size)
        self.encoder_linear = nn.Linear(cfg.human_human_edge_input_size, cfg.human_human_edge_embedding_size)
        _init_gru(self.gru)


class _NodeRNN(nn.Module):
    """Parameter holder of HumanNodeRNN (srnn_model.py:109-173)."""

    def __init__(self, cfg):
        super().__init__()
        self.gru = nn.GRU(cfg.human_node_embedding_size * 2, cfg.human_node_rnn_size)
        self.encoder_linear = nn.Linear(cfg.human_node_input_size, cfg.human_node_embedding_size)
        self.edge_embed = nn.Linear(cfg.human_human_edge_rnn_size, cfg.human_node_embedding_size)  # unused, kept for state_dict
        self.edge_attention_embed = nn.Linear(cfg.human_human_edge_rnn_size * 2, cfg.human_node_embedding_size)
        self.output_linear = nn.Linear(cfg.human_node_rnn_size, cfg.human_node_output_size)
        _init_gru(self.gru)


class _Attention(nn.Module):
    """Parameter holder of EdgeAttention (srnn_model.py:218-339)."""

    def __init__(self, cfg):
        super().__init__()
        self.temporal_edge_layer = nn.ModuleList([nn.Linear(cfg.human_human_edge_rnn_size, cfg.attention_size)])
        self.spatial_edge_layer = nn.ModuleList([nn.Linear(cfg.human_human_edge_rnn_size, cfg.attention_size)])


def _init_gru(gru):
    for name, param in gru.named_parameters():
        if "bias" in name:
            nn.init.constant_(param, 0)
        elif "weight" in name:
            nn.init.orthogonal_(param)


class SRNN(nn.Module):
    """Module tree (names = state_dict keys) of srnn_model.py:342-407."""

    def __init__(self, obs_space_dict, config, infer=False):
        super().__init__()
        self.infer = infer
        self.is_recurrent = True
        self.config = config
        self.human_num = config.sim.human_num
        self.seq_length = config.ppo.num_steps
        self.nenv = config.training.num_processes
        self.nminibatch = config.ppo.num_mini_batch
        c = config.SRNN
        if (c.human_node_rnn_size, c.human_human_edge_rnn_size, c.human_node_output_size, c.human_node_embedding_size,
                c.human_human_edge_embedding_size, c.attention_size, c.human_node_input_size,
                c.human_human_edge_input_size) != (128, 256, 256, 64, 64, 64, 3, 2):
            raise NotImplementedError("the CUDA forward is specialised for the reference's SRNN sizes (config.py:178-193)")
        self.human_node_rnn_size = c.human_node_rnn_size
        self.human_human_edge_rnn_size = c.human_human_edge_rnn_size
        self.output_size = c.human_node_output_size
        self.humanNodeRNN = _NodeRNN(c)
        self.humanhumanEdgeRNN_spatial = _EdgeRNN(c)
        self.humanhumanEdgeRNN_temporal = _EdgeRNN(c)
        self.attn = _Attention(c)
        g = math.sqrt(2)
        hid = self.output_size
        self.actor = nn.Sequential(_orthogonal(nn.Linear(hid, hid), g), nn.Tanh(), _orthogonal(nn.Linear(hid, hid), g), nn.Tanh())
        self.critic = nn.Sequential(_orthogonal(nn.Linear(hid, hid), g), nn.Tanh(), _orthogonal(nn.Linear(hid, hid), g), nn.Tanh())
        self.critic_linear = _orthogonal(nn.Linear(hid, 1), g)
        self.robot_linear = _orthogonal(nn.Linear(7, 3), g)
        self.human_node_final_linear = _orthogonal(nn.Linear(hid, 2), g)  # unused, kept for state_dict
        self.num_edges = self.human_num + 1


class _AddBias(nn.Module):
    def __init__(self, bias):
        super().__init__()
        self._bias = nn.Parameter(bias.unsqueeze(1))


class DiagGaussian(nn.Module):
    """distributions.py:74-94: fc_mean + state-independent logstd."""

    def __init__(self, num_inputs, num_outputs):
        super().__init__()
        self.fc_mean = _orthogonal(nn.Linear(num_inputs, num_outputs))
        self.logstd = _AddBias(torch.zeros(num_outputs))

    def std(self):
        return self.logstd._bias.t().view(1, -1).exp()


def _normal_log_prob(action, mean, std):
    var = std * std
    return (-((action - mean) ** 2) / (2 * var) - std.log() - math.log(math.sqrt(2 * math.pi))).sum(-1, keepdim=True)


class Policy(nn.Module):
    def __init__(self, obs_shape, action_space, base=None, base_kwargs=None):
        super().__init__()
        if base != "srnn":
            raise NotImplementedError("only base='srnn' is on the hot path (convgru is out of scope, SURVEY section 2 row 18)")
        self.base = SRNN(obs_shape, base_kwargs)
        self.srnn = True
        if action_space.__class__.__name__ != "Box":
            raise NotImplementedError("only Box action spaces are supported")
        self.dist = DiagGaussian(self.base.output_size, action_space.shape[0])
        self.precision = "bf16x3"   # contraction precision of the CUDA forward: "fp32" | "bf16x3" | "bf16"
        self.sequence_impl = "batched"   # evaluate_actions: "batched" (whole [T, N] chunk at once) | "per_step"
        self._handle = None
        self._weights_key = None
        self._workspace = None
        self.gpu_launches = 0

    @property
    def is_recurrent(self):
        return self.base.is_recurrent

    def forward(self, inputs, rnn_hxs, masks):
        raise NotImplementedError

    # ------------------------------------------------------------------ CUDA forward plumbing
    def _weight_tensors(self):
        sd = dict(self.named_parameters())
        return [sd[abi.DSRNN_STATE_DICT_KEYS[f]] for f in abi.DSRNN_WEIGHT_FIELDS]

    _TRANSIENT = ("_handle", "_weights_key", "_workspace", "_workspace_key", "_param_list", "_gauss_cache", "_refill_engine", "_spare_out")

    def __getstate__(self):                     # copy.deepcopy / torch.save(policy): library handles and caches stay behind
        state = dict(self.__dict__)
        for k in self._TRANSIENT:
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.__dict__.update(_handle=None, _weights_key=None, _workspace=None)

    def _apply(self, fn, *args, **kwargs):      # .to() / .cuda() / .float(): forget the cached parameter list
        self.__dict__.pop("_param_list", None)
        return super()._apply(fn, *args, **kwargs)

    def _ensure_handle(self, device):
        # called once per forward, right after the host synchronisation of the previous env step in a train.py-style
        # loop: keep it to ~40 attribute reads (Parameter objects are stable; their storage / version are not)
        tensors = self.__dict__.get("_param_list")
        if tensors is None:
            tensors = self._weight_tensors()
            self.__dict__["_param_list"] = tensors
        key = (device, tuple([(t.data_ptr(), t._version) for t in tensors]))
        lib = _lib.load()
        if self._handle is not None and key == self._weights_key:
            return lib
        for t in tensors:
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                raise _lib.CrowdNavLibraryError("Policy parameters must be contiguous float32 on %s (call .to(device))" % device)
        w = abi.CnDsrnnWeights(*[_ptr(t) for t in tensors])
        stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is None:
            handle = C.c_void_p()
            _lib.check(lib.cn_dsrnn_create(C.byref(w), index, stream, C.byref(handle)), "cn_dsrnn_create")
            self._handle = handle
            hook = self.__dict__.get("_refill_engine")
            if hook is not None:
                _lib.check(lib.cn_dsrnn_set_refill_env(handle, hook.handle), "cn_dsrnn_set_refill_env")
        else:
            _lib.check(lib.cn_dsrnn_update_weights(self._handle, C.byref(w), stream), "cn_dsrnn_update_weights")
        self._weights_key = key
        return lib

    def invalidate_weights(self):
        """Force the next forward to re-pack the weights (call after editing parameters through `.data` or any other
        path that does not bump their version counter)."""
        self._weights_key = None
        self.__dict__.pop("_gauss_cache", None)

    def __del__(self):
        try:  # best effort: at interpreter shutdown torch / ctypes may already be torn down
            handle = self.__dict__.get("_handle")
            if handle is not None:
                self.__dict__["_handle"] = None
                _lib.load().cn_dsrnn_destroy(handle)
        except Exception:  # noqa: BLE001
            pass

    def start_refill_of(self, engine):
        """Let every forward start `engine`'s spare-episode refill beside its attention kernel (pair with
        `engine.step(..., defer_refill=True)`; used by rollout.GraphedRollout).  `engine=None` clears the hook."""
        self.__dict__["_refill_engine"] = engine
        if engine is not None:
            engine.__dict__.setdefault("_refill_policies", []).append(weakref.ref(self))
        if self._handle is not None:
            _lib.check(_lib.load().cn_dsrnn_set_refill_env(self._handle, engine.handle if engine is not None else None),
                       "cn_dsrnn_set_refill_env")

    def set_edge_event(self, event):
        """Have the next forwards record `event` (a torch.cuda.Event, None clears it) right behind their edge-GRU stage, so
        that another stream can start independent work beside the rest of the forward (rollout.PipelinedRollout)."""
        if self._handle is None:
            raise _lib.CrowdNavLibraryError("set_edge_event needs the library handle: run one forward first")
        _lib.check(_lib.load().cn_dsrnn_set_edge_event(self._handle, C.c_void_p(event.cuda_event) if event is not None else None),
                   "cn_dsrnn_set_edge_event")

    def set_edge_image(self, image_in=None, image_out=None):
        """Resident split-bf16 image of the edge hidden state for the next forwards (rollout.GraphedRollout): `image_out` =
        (hi, lo) bfloat16 tensors [N * (H + 1), 256] the forward fills for its h_edge output; `image_in` = the pair a previous
        forward filled for exactly the h_edge tensor passed in (masks must be 0 / 1).  None switches either half off."""
        if self._handle is None:
            raise _lib.CrowdNavLibraryError("set_edge_image needs the library handle: run one forward first")
        ptrs = []
        for pair in (image_in, image_out):
            for t in (pair if pair is not None else (None, None)):
                if t is not None and (t.dtype is not torch.bfloat16 or not t.is_contiguous() or not t.is_cuda):
                    raise ValueError("edge images are contiguous CUDA bfloat16 tensors")
                ptrs.append(_ptr(t))
        _lib.check(_lib.load().cn_dsrnn_set_edge_image(self._handle, *ptrs), "cn_dsrnn_set_edge_image")

    def cuda_forward(self, inputs, rnn_hxs, masks, need_features=True, out=None, workspace=None):
        """One rollout-step forward on the GPU. Returns (value[N,1], mean[N,2], features[N,256]|None, h_node, h_edge).
        `out` = dict of preallocated outputs (h_node, h_edge, value, mean) for callers that need static addresses
        (CUDA-graph capture, rollout.py); `workspace` = a caller-owned uint8 scratch tensor of at least
        `cn_dsrnn_workspace_bytes(N, H)` bytes for forwards that run concurrently on several streams."""
        rn, te, se = inputs["robot_node"], inputs["temporal_edges"], inputs["spatial_edges"]
        device = se.device
        if device.type != "cuda":
            raise _lib.CrowdNavLibraryError("Policy.act/get_value run on a B200 only; inputs are on %s (no CPU fallback)" % device)
        lib = self._ensure_handle(device)
        N, H = se.shape[0], se.shape[1]
        f = lambda t: t if (t.dtype is torch.float32 and t.is_contiguous()) else t.detach().to(dtype=torch.float32).contiguous()
        rn, te, se = f(rn), f(te), f(se)
        hn, he, mk = f(rnn_hxs["human_node_rnn"]), f(rnn_hxs["human_human_edge_rnn"]), f(masks)
        if hn.numel() != N * 128 or he.numel() != N * (H + 1) * 256 or mk.numel() != N or rn.numel() != N * 7:
            raise ValueError("inconsistent batch shapes for the DS-RNN forward")
        opts = dict(dtype=torch.float32, device=device)
        if out is None:
            # fresh output tensors every call (the reference returns new tensors); they were allocated at the END of the
            # previous call, while the GPU was busy, so the allocator is off the critical path of a train.py-style loop
            stream = _lib.raw_stream(device.index if device.index is not None else torch.cuda.current_device())
            spare = self.__dict__.pop("_spare_out", None)
            if spare is not None and spare[0] == (N, H, device, stream.value):      # same stream: the allocator's own ordering holds
                hn_out, he_out, value, mean = spare[1]
            else:
                hn_out, he_out, value, mean = self._alloc_out(N, H, opts)
        else:
            hn_out, he_out, value, mean = out["h_node"], out["h_edge"], out["value"], out["mean"]
            for t_, numel in ((hn_out, N * 128), (he_out, N * (H + 1) * 256), (value, N), (mean, 2 * N)):
                if t_.numel() != numel or t_.dtype != torch.float32 or t_.device != device or not t_.is_contiguous():
                    raise ValueError("preallocated forward outputs have the wrong shape / dtype / device")
        feat = torch.empty(N, 256, **opts) if need_features else None
        if workspace is None:
            ws_key = (N, H, device)
            if self.__dict__.get("_workspace_key") != ws_key:
                nbytes = lib.cn_dsrnn_workspace_bytes(N, H)
                if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != device:
                    self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=device)
                self.__dict__["_workspace_key"] = ws_key
            workspace = self._workspace
        io = abi.CnDsrnnIO(_ptr(rn), _ptr(te), _ptr(se), _ptr(hn), _ptr(he), _ptr(mk), _ptr(hn_out), _ptr(he_out),
                           _ptr(value), _ptr(mean), _ptr(feat))
        if out is not None:
            stream = _lib.raw_stream(device.index if device.index is not None else torch.cuda.current_device())
        _lib.check(lib.cn_dsrnn_forward(self._handle, N, H, C.byref(io), abi.PRECISIONS[self.precision],
                                        _ptr(workspace), workspace.numel(), stream), "cn_dsrnn_forward")
        self.gpu_launches += lib.cn_dsrnn_last_launches(self._handle)
        if out is None:
            self.__dict__["_spare_out"] = ((N, H, device, stream.value), self._alloc_out(N, H, opts))
        return value, mean, feat, hn_out, he_out

    @staticmethod
    def _alloc_out(N, H, opts):
        return (torch.empty(N, 1, 128, **opts), torch.empty(N, H + 1, 256, **opts), torch.empty(N, 1, **opts), torch.empty(N, 2, **opts))

    # ------------------------------------------------------------------ reference API
    def act(self, inputs, rnn_hxs, masks, deterministic=False):
        with torch.no_grad():
            value, mean, _, hn, he = self.cuda_forward(inputs, rnn_hxs, masks, need_features=False)
            # DiagGaussian sample + log-prob (distributions.py:36-45, 85-94) in as few launches as possible -- in a
            # train.py-style loop this runs on the host's critical path right after the env step's synchronisation:
            # log N(a; mean, std) = -0.5 sum(eps^2) - sum(logstd) - A/2 log(2 pi) with eps = (a - mean) / std
            bias = self.dist.logstd._bias
            cached = self.__dict__.get("_gauss_cache")
            if cached is None or cached[0] != (bias.data_ptr(), bias._version):      # recomputed after an optimiser step
                logstd = bias.detach().view(1, -1)
                const = -logstd.sum() - 0.5 * mean.shape[-1] * math.log(2 * math.pi)      # 0-dim, stays on the device
                cached = ((bias.data_ptr(), bias._version), const, logstd.exp())
                self.__dict__["_gauss_cache"] = cached
            _, const, std = cached
            if deterministic:
                action = mean
                log_probs = const.expand(mean.shape[0], 1)
            else:
                eps = torch.randn_like(mean)
                action = torch.addcmul(mean, eps, std)
                log_probs = eps.square().sum(-1, keepdim=True).mul_(-0.5).add_(const)
        rnn_hxs["human_node_rnn"] = hn            # the reference mutates the dict it was given (srnn_model.py:481-491)
        rnn_hxs["human_human_edge_rnn"] = he
        return value, action, log_probs, rnn_hxs

    def get_value(self, inputs, rnn_hxs, masks):
        with torch.no_grad():
            value, _, _, hn, he = self.cuda_forward(inputs, rnn_hxs, masks, need_features=False)
        rnn_hxs["human_node_rnn"] = hn
        rnn_hxs["human_human_edge_rnn"] = he
        return value

    def evaluate_actions(self, inputs, rnn_hxs, masks, action):
        """Training-time evaluation over a [T*N, ...] rollout chunk (model.py:96-104; srnn_model.py:53-104)."""
        value, feat, rnn_hxs = self._torch_sequence_forward(inputs, rnn_hxs, masks)
        mean = self.dist.fc_mean(feat)
        std = self.dist.std().expand_as(mean)
        log_probs = _normal_log_prob(action, mean, std)
        # distributions.py:40 spells FixedNormal's override "entrop", so model.py:102 gets torch's element-wise Normal
        # entropy and its .mean() averages over the action dimensions too (not summed over them); kept as is.
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + std.log()).mean()
        return value, log_probs, entropy, rnn_hxs

    # ------------------------------------------------------------------ differentiable torch restatement (training)
    def _torch_step(self, rn, te, se, h_node, h_edge, m):
        b = self.base
        N, H = se.shape[0], se.shape[1]

        def gru(mod, x, h):   # one GRU step; ATen's fused cell (two GEMMs + one gate kernel, with autograd)
            return torch.gru_cell(x, h, mod.weight_ih_l0, mod.weight_hh_l0, mod.bias_ih_l0, mod.bias_hh_l0)

        h_edge = h_edge * m.view(N, 1, 1)
        h_node = h_node * m.view(N, 1)
        o_t = gru(b.humanhumanEdgeRNN_temporal.gru, torch.relu(b.humanhumanEdgeRNN_temporal.encoder_linear(te)), h_edge[:, 0])
        o_s = gru(b.humanhumanEdgeRNN_spatial.gru, torch.relu(b.humanhumanEdgeRNN_spatial.encoder_linear(se.reshape(N * H, 2))),
                  h_edge[:, 1:].reshape(N * H, 256)).view(N, H, 256)
        q = b.attn.temporal_edge_layer[0](o_t)
        k = b.attn.spatial_edge_layer[0](o_s)
        alpha = torch.softmax((q.unsqueeze(1) * k).sum(-1) * (H / math.sqrt(64.0)), dim=-1)
        c = (alpha.unsqueeze(-1) * o_s).sum(1)
        enc = torch.relu(b.humanNodeRNN.encoder_linear(b.robot_linear(rn)))
        emb = torch.relu(b.humanNodeRNN.edge_attention_embed(torch.cat([o_t, c], -1)))
        h_n = gru(b.humanNodeRNN.gru, torch.cat([enc, emb], -1), h_node)
        y = b.humanNodeRNN.output_linear(h_n)
        return b.critic_linear(b.critic(y)), b.actor(y), h_n, torch.cat([o_t.unsqueeze(1), o_s], 1)

    def _torch_sequence_forward(self, inputs, rnn_hxs, masks):
        """The whole [T, N] chunk at once.  The edge GRUs do not depend on the node RNN (srnn_model.py:464-480), so their
        T steps run first as one `_MaskedGruSequence` each; the attention, the encoders and the heads have no recurrence
        and are evaluated for all T*N samples in one batch; the node GRU is a second (small) masked sequence."""
        impl = getattr(self, "sequence_impl", "batched")
        if impl == "per_step":
            return self._torch_sequence_forward_per_step(inputs, rnn_hxs, masks)
        if impl == "native":
            return self._native_sequence_forward(inputs, rnn_hxs, masks)
        b = self.base
        se = inputs["spatial_edges"]
        H = se.shape[1]
        N = rnn_hxs["human_node_rnn"].shape[0]
        T = se.shape[0] // N
        rn = inputs["robot_node"].reshape(T * N, 7)
        te = inputs["temporal_edges"].reshape(T, N, 2)
        se = se.reshape(T, N * H, 2)
        mk = masks.reshape(T, N, 1)
        h_node = rnn_hxs["human_node_rnn"].reshape(N, 128)
        h_edge = rnn_hxs["human_human_edge_rnn"].reshape(N, H + 1, 256)

        def seq(mod, x, h0, m):
            return _MaskedGruSequence.apply(x, h0, m, mod.weight_ih_l0, mod.weight_hh_l0, mod.bias_ih_l0, mod.bias_hh_l0)

        et, es = b.humanhumanEdgeRNN_temporal, b.humanhumanEdgeRNN_spatial
        o_t = seq(et.gru, torch.relu(et.encoder_linear(te)), h_edge[:, 0], mk)                               # [T, N, 256]
        mk_rows = mk.expand(T, N, H).reshape(T, N * H, 1)
        o_s = seq(es.gru, torch.relu(es.encoder_linear(se)), h_edge[:, 1:].reshape(N * H, 256), mk_rows)    # [T, N*H, 256]
        o_t2, o_s3 = o_t.view(T * N, 256), o_s.view(T * N, H, 256)
        q = b.attn.temporal_edge_layer[0](o_t2)
        k = b.attn.spatial_edge_layer[0](o_s3)
        alpha = torch.softmax(torch.bmm(k, q.unsqueeze(-1)).squeeze(-1) * (H / math.sqrt(64.0)), dim=-1)
        c = torch.bmm(alpha.unsqueeze(1), o_s3).squeeze(1)
        enc = torch.relu(b.humanNodeRNN.encoder_linear(b.robot_linear(rn)))
        emb = torch.relu(b.humanNodeRNN.edge_attention_embed(torch.cat([o_t2, c], -1)))
        h_n = seq(b.humanNodeRNN.gru, torch.cat([enc, emb], -1).view(T, N, -1), h_node, mk)                 # [T, N, 128]
        y = b.humanNodeRNN.output_linear(h_n.view(T * N, 128))
        rnn_hxs["human_node_rnn"] = h_n[-1].unsqueeze(1)
        rnn_hxs["human_human_edge_rnn"] = torch.cat([o_t[-1].unsqueeze(1), o_s[-1].view(N, H, 256)], 1)
        return b.critic_linear(b.critic(y)), b.actor(y), rnn_hxs

    def _native_sequence_forward(self, inputs, rnn_hxs, masks):
        """The batched form with every contraction on the library's own tensor-core kernels (native.py): the two edge GRUs as
        one `EdgeGruSequence` (the rollout's tcgen05 edge kernel in training mode + its backward), every linear layer and
        the node GRU's products on cn_gemm_bf16x3.  The attention's key projection is folded into the query exactly as in
        the rollout kernel: q.(W_s o_i + b_s) = (W_s^T q).o_i + q.b_s, and the softmax ignores the per-sample constant."""
        global SEQUENCE_GEMM
        b = self.base
        se = inputs["spatial_edges"]
        if se.device.type != "cuda":
            raise _lib.CrowdNavLibraryError("sequence_impl='native' runs on a B200 only (no CPU fallback); use 'batched' on the CPU")
        H = se.shape[1]
        N = rnn_hxs["human_node_rnn"].shape[0]
        T = se.shape[0] // N
        S = N * H
        rn = inputs["robot_node"].reshape(T * N, 7)
        mk = masks.reshape(T, N)
        h_node = rnn_hxs["human_node_rnn"].reshape(N, 128)
        h_edge = rnn_hxs["human_human_edge_rnn"].reshape(N, H + 1, 256)
        h0 = torch.cat([h_edge[:, 1:].reshape(S, 256), h_edge[:, 0]], 0)
        es, et = b.humanhumanEdgeRNN_spatial, b.humanhumanEdgeRNN_temporal
        edge_params = [p for m in (es, et) for p in (m.encoder_linear.weight, m.encoder_linear.bias, m.gru.weight_ih_l0,
                                                     m.gru.weight_hh_l0, m.gru.bias_ih_l0, m.gru.bias_hh_l0)]
        o_s, o_t = native.EdgeGruSequence.apply(self, se.reshape(T, S, 2), inputs["temporal_edges"].reshape(T, N, 2), h0, mk, *edge_params)
        o_s, o_t = o_s.view(T, S, 256), o_t.view(T, N, 256)
        o_t2, o_s3 = o_t.reshape(T * N, 256), o_s.view(T * N, H, 256)
        q = native.linear(o_t2, b.attn.temporal_edge_layer[0])                                   # [TN, 64]
        qt = native.matmul_nt(q, b.attn.spatial_edge_layer[0].weight.t())                        # [TN, 256] = q W_s
        const = (q * b.attn.spatial_edge_layer[0].bias).sum(-1)                                  # q.b_s (no effect on the softmax)
        c = native.AttentionMix.apply(o_s3, qt, const, H / math.sqrt(64.0))
        enc = torch.relu(b.humanNodeRNN.encoder_linear(b.robot_linear(rn)))                      # K = 7 / 3: CUDA cores
        emb = native.linear(torch.cat([o_t2, c], -1), b.humanNodeRNN.edge_attention_embed, "relu")
        prev, SEQUENCE_GEMM = SEQUENCE_GEMM, "native"
        try:
            g = b.humanNodeRNN.gru
            h_n = _MaskedGruSequence.apply(torch.cat([enc, emb], -1).view(T, N, -1), h_node, mk.reshape(T, N, 1),
                                           g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0)     # [T, N, 128]
        finally:
            SEQUENCE_GEMM = prev
        y = native.linear(h_n.view(T * N, 128), b.humanNodeRNN.output_linear)
        actor = native.linear(native.linear(y, b.actor[0], "tanh"), b.actor[2], "tanh")
        critic = native.linear(native.linear(y, b.critic[0], "tanh"), b.critic[2], "tanh")
        rnn_hxs["human_node_rnn"] = h_n[-1].unsqueeze(1)
        rnn_hxs["human_human_edge_rnn"] = torch.cat([o_t[-1].unsqueeze(1), o_s[-1].view(N, H, 256)], 1)
        return b.critic_linear(critic), actor, rnn_hxs

    def _torch_sequence_forward_per_step(self, inputs, rnn_hxs, masks):
        """One autograd graph per rollout step (the first form of this path; kept as the cross-check of the batched one)."""
        se = inputs["spatial_edges"]
        H = se.shape[1]
        N = rnn_hxs["human_node_rnn"].shape[0]
        T = se.shape[0] // N
        rn = inputs["robot_node"].reshape(T, N, 7)
        te = inputs["temporal_edges"].reshape(T, N, 2)
        se = se.reshape(T, N, H, 2)
        mk = masks.reshape(T, N)
        h_node = rnn_hxs["human_node_rnn"].reshape(N, 128)
        h_edge = rnn_hxs["human_human_edge_rnn"].reshape(N, H + 1, 256)
        values, feats = [], []
        for t in range(T):
            v, f, h_node, h_edge = self._torch_step(rn[t], te[t], se[t], h_node, h_edge, mk[t])
            values.append(v)
            feats.append(f)
        rnn_hxs["human_node_rnn"] = h_node.unsqueeze(1)
        rnn_hxs["human_human_edge_rnn"] = h_edge
        return torch.cat(values, 0).view(-1, 1), torch.cat(feats, 0).view(-1, self.base.output_size), rnn_hxs

"""Helpers for the -m gpu parity tests: run the CUDA path through the C ABI (CrowdEngine) and the C oracle
on identical inputs and return both as numpy dicts in the canonical layout."""
import numpy as np
import torch

from crowdnav_dsrnn_b200 import abi
from crowdnav_dsrnn_b200.engine import CrowdEngine
from oracle import crowd_oracle

OUT_FIELDS = ("robot_node", "temporal_edges", "spatial_edges", "visible_mask", "reward", "done", "event", "scenario",
              "info", "episode_return", "episode_length", "goal_changed")
STATE_FIELDS = ("robot", "humans", "belief", "extras", "counters", "episode_return", "groups")
INT_FIELDS = ("visible_mask", "done", "event", "scenario", "episode_length", "goal_changed", "counters")


def buf_to_numpy(buf):
    out = {}
    for f in OUT_FIELDS:
        a = getattr(buf, f).detach().cpu().numpy()
        if f in ("visible_mask", "goal_changed"):
            a = a.astype(np.int64) & 0xFFFFFFFF
        out[f] = a
    return out


def oracle_out_to_numpy(o):
    out = o.as_dict()
    for f in ("visible_mask", "goal_changed"):
        out[f] = out[f].astype(np.int64)
    return out


def state_to_numpy(st):
    return {k: v.detach().cpu().numpy() for k, v in st.items()}


def make_engine(cfg_obj, n, **kw):
    return CrowdEngine(cfg_obj, n, torch.device("cuda:0"), phase="train", **kw)


def oracle_state_from(inp, n, H):
    st = crowd_oracle.OracleState(n, H)
    for f in ("robot", "humans", "belief", "extras", "counters"):
        getattr(st, f)[...] = inp[f]
    if "episode_return" in inp:
        st.episode_return[...] = inp["episode_return"]
    if "groups" in inp:
        st.groups[...] = inp["groups"]
    return st


def compare(gpu_out, gpu_state, or_out, or_state, float_tol=2e-6, where=""):
    """Integer / flag outputs bit-exact; floating point within float_tol. Returns a report dict."""
    rep = {}
    for f in OUT_FIELDS:
        a, b = gpu_out[f], or_out[f]
        if f in INT_FIELDS:
            assert np.array_equal(a, b), "%s %s: %d mismatches" % (where, f, int((a != b).sum()))
        else:
            both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
            err = np.where(both_inf, 0.0, np.abs(a.astype(np.float64) - b.astype(np.float64)))
            rep[f] = float(err.max())
            assert rep[f] <= float_tol, "%s %s max err %g" % (where, f, rep[f])
    for f in STATE_FIELDS:
        a, b = gpu_state[f], getattr(or_state, f)
        if f in INT_FIELDS:
            assert np.array_equal(a, b), "%s state.%s: %d mismatches" % (where, f, int((a != b).sum()))
        else:
            rep["state." + f] = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())
            assert rep["state." + f] <= float_tol, "%s state.%s max err %g" % (where, f, rep["state." + f])
    rep["bit_exact_humans"] = float(np.mean(gpu_state["humans"] == or_state.humans))
    return rep

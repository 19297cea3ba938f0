"""TEST INFRASTRUCTURE -- numpy/ctypes binding of oracle/_build/libcrowd_oracle.so
(the plain-C CPU restatement of the crowd step, oracle/crowd_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.  The product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from crowdnav_dsrnn_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile the C oracle with gcc (oracle/Makefile)."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libcrowd_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_step.restype = C.c_int
        L.oracle_reset.restype = C.c_int
        L.oracle_observe.restype = C.c_int
        L.oracle_orca_one.restype = C.c_int
        _LIB = L
    return _LIB


class OrStateView(C.Structure):
    _fields_ = abi.CnStateView._fields_


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleState:
    """Canonical env state in host numpy arrays (field order of CnStateView)."""

    def __init__(self, n_envs, human_num):
        self.n, self.h = n_envs, human_num
        self.robot = np.zeros((n_envs, 9), np.float32)
        self.humans = np.zeros((n_envs, human_num, 9), np.float32)
        self.belief = np.zeros((n_envs, human_num, 5), np.float32)
        self.extras = np.zeros((n_envs, 4), np.float32)
        self.counters = np.zeros((n_envs, 4), np.int32)
        self.episode_return = np.zeros((n_envs,), np.float32)
        self.groups = np.zeros((n_envs, abi.MAX_GROUPS, 4), np.float32)

    FIELDS = ("robot", "humans", "belief", "extras", "counters", "episode_return", "groups")

    def view(self):
        return OrStateView(*[_ptr(getattr(self, f)) for f in self.FIELDS])

    def copy(self):
        o = OracleState(self.n, self.h)
        for f in self.FIELDS:
            getattr(o, f)[...] = getattr(self, f)
        return o

    def as_dict(self):
        return {f: getattr(self, f) for f in self.FIELDS}


class StepOut:
    def __init__(self, n_envs, human_num):
        n, h = n_envs, human_num
        self.robot_node = np.zeros((n, 1, 7), np.float32)
        self.temporal_edges = np.zeros((n, 1, 2), np.float32)
        self.spatial_edges = np.zeros((n, h, 2), np.float32)
        self.visible_mask = np.zeros((n,), np.uint32)
        self.reward = np.zeros((n,), np.float32)
        self.done = np.zeros((n,), np.uint8)
        self.event = np.zeros((n,), np.int32)
        self.scenario = np.zeros((n,), np.int32)
        self.info = np.zeros((n, abi.INFO_DIM), np.float32)
        self.episode_return = np.zeros((n,), np.float32)
        self.episode_length = np.zeros((n,), np.int32)
        self.goal_changed = np.zeros((n,), np.uint32)

    def obs_struct(self):
        return abi.CnObsOut(_ptr(self.robot_node), _ptr(self.temporal_edges), _ptr(self.spatial_edges),
                            _ptr(self.visible_mask))

    def struct(self):
        return abi.CnStepOut(self.obs_struct(), _ptr(self.reward), _ptr(self.done), _ptr(self.event),
                             _ptr(self.scenario), _ptr(self.info), _ptr(self.episode_return),
                             _ptr(self.episode_length), _ptr(self.goal_changed))

    FIELDS = ("robot_node", "temporal_edges", "spatial_edges", "visible_mask", "reward", "done", "event",
              "scenario", "info", "episode_return", "episode_length", "goal_changed")

    def as_dict(self):
        return {f: getattr(self, f) for f in self.FIELDS}


def step(cfg, state, action, auto_reset=False, n_threads=1):
    """One oracle step on every env of `state` (mutated in place). Returns StepOut."""
    action = np.ascontiguousarray(action, np.float32).reshape(state.n, 2)
    out = StepOut(state.n, state.h)
    so, sv = out.struct(), state.view()
    rc = lib().oracle_step(C.byref(cfg), C.c_int(state.n), C.byref(sv), _ptr(action), C.byref(so),
                           C.c_int(int(auto_reset)), C.c_int(n_threads))
    if rc != 0:
        raise RuntimeError("oracle_step failed (%d)" % rc)
    return out


def reset(cfg, state, mask=None, n_threads=1):
    out = StepOut(state.n, state.h)
    ob, sv = out.obs_struct(), state.view()
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
    rc = lib().oracle_reset(C.byref(cfg), C.c_int(state.n), C.byref(sv), _ptr(mask), C.byref(ob), C.c_int(n_threads))
    if rc != 0:
        raise RuntimeError("oracle_reset failed (%d)" % rc)
    return out


def observe(cfg, state):
    out = StepOut(state.n, state.h)
    ob, sv = out.obs_struct(), state.view()
    rc = lib().oracle_observe(C.byref(cfg), C.c_int(state.n), C.byref(sv), C.byref(ob), C.c_int(1))
    if rc != 0:
        raise RuntimeError("oracle_observe failed (%d)" % rc)
    return out


def philox(key, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().oracle_philox(C.c_uint64(key), C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(c3), out)
    return [int(x) for x in out]


def orca_one(cfg, self9, others5):
    self9 = np.ascontiguousarray(self9, np.float32)
    others5 = np.ascontiguousarray(others5, np.float32).reshape(-1, 5)
    out = np.zeros(2, np.float32)
    rc = lib().oracle_orca_one(C.byref(cfg), _ptr(self9), C.c_int(len(others5)), _ptr(others5), _ptr(out))
    if rc != 0:
        raise RuntimeError("oracle_orca_one failed")
    return out

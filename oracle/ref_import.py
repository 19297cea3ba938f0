"""TEST INFRASTRUCTURE -- makes the reference's own Python importable in THIS
container (it does not exist on the GPU box; nothing under tests -m gpu,
smoke() or bench.py may call this).

The reference needs gym, OpenAI baselines, matplotlib, shapely and rvo2, none
of which are installed or installable here.  This
module registers:

* inert stand-ins for `gym`, `baselines.*`, `matplotlib.*` (only what import
  time and the CrowdSimDict / Policy code paths touch);
* exact-geometry stand-ins for `shapely` (oracle/refshim/shapely_standin.py);
* `rvo2` = ctypes binding of oracle/rvo2_port.c (restated RVO2).

Declared oracle patches: P2 `numpy.bool` alias.  P1 (unicycle
crash in calc_reward) and P3 (state injection) live in oracle/ref_harness.py.
"""
import os
import sys
import types
from contextlib import contextmanager

import numpy as np

REFERENCE_ROOT = os.environ.get("CROWDNAV_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "crowd_sim"))


class _Inert(types.ModuleType):
    """Module whose unknown attributes are inert callables/objects (dunder lookups still fail)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        obj = _InertObj(f"{self.__name__}.{name}")
        setattr(self, name, obj)
        return obj


class _InertObj:
    def __init__(self, name="inert"):
        self._name = name

    def __call__(self, *a, **k):
        return _InertObj(self._name + "()")

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _InertObj(self._name + "." + name)

    def __iter__(self):
        return iter(())


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


# --------------------------------------------------------------------------- gym
class _Space:
    pass


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)


class Dict(_Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)


class Env:
    metadata = {}
    observation_space = None
    action_space = None

    @property
    def unwrapped(self):
        return self

    def seed(self, seed=None):
        return [seed]

    def close(self):
        pass

    def render(self, mode="human"):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = env.observation_space
        self.action_space = env.action_space

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kw):
        return self.env.reset(**kw)


class ObservationWrapper(Wrapper):
    pass


_REGISTRY = {}


def _register(id, entry_point, **kw):
    _REGISTRY[id] = entry_point


def _make(id, **kw):
    import importlib

    mod, _, cls = _REGISTRY[id].partition(":")
    return getattr(importlib.import_module(mod), cls)(**kw)


# --------------------------------------------------------------------- baselines
class VecEnv:
    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close_extras(self):
        pass

    def close(self):
        self.close_extras()

    def render(self, mode="human"):
        pass

    @property
    def unwrapped(self):
        return self


class VecEnvWrapper(VecEnv):
    def __init__(self, venv, observation_space=None, action_space=None):
        self.venv = venv
        super().__init__(venv.num_envs, observation_space or venv.observation_space,
                         action_space or venv.action_space)

    def step_async(self, actions):
        self.venv.step_async(actions)

    def close(self):
        return self.venv.close()

    def render(self, mode="human"):
        return self.venv.render(mode=mode)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.venv, name)


def obs_space_info(obs_space):
    subspaces = obs_space.spaces if isinstance(obs_space, Dict) else {None: obs_space}
    keys, shapes, dtypes = [], {}, {}
    for key, box in subspaces.items():
        keys.append(key)
        shapes[key] = box.shape
        dtypes[key] = box.dtype
    return keys, shapes, dtypes


def obs_to_dict(obs):
    return obs if isinstance(obs, dict) else {None: obs}


def dict_to_obs(obs_dict):
    return obs_dict[None] if set(obs_dict.keys()) == {None} else obs_dict


class DummyVecEnv(VecEnv):
    """Functional restatement (sequential envs, auto-reset on done)."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        env = self.envs[0]
        super().__init__(len(env_fns), env.observation_space, env.action_space)
        self.keys, _, _ = obs_space_info(env.observation_space)
        self.actions = None

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        obs_l, rews, dones, infos = [], [], [], []
        for e, env in enumerate(self.envs):
            ob, rew, done, info = env.step(self.actions[e])
            if done:
                ob = env.reset()
            obs_l.append(ob)
            rews.append(rew)
            dones.append(done)
            infos.append(info)
        return self._stack(obs_l), np.array(rews, dtype=np.float32), np.array(dones), infos

    def reset(self):
        return self._stack([env.reset() for env in self.envs])

    def _stack(self, obs_l):
        if isinstance(obs_l[0], dict):
            return {k: np.stack([np.asarray(o[k], dtype=np.float32) for o in obs_l]) for k in obs_l[0]}
        return np.stack(obs_l)


class Monitor(Wrapper):
    """Functional restatement of baselines.bench.Monitor: adds info['episode'] on done."""

    def __init__(self, env, filename=None, allow_early_resets=False, **kw):
        super().__init__(env)
        self.rewards = []

    def reset(self, **kw):
        self.rewards = []
        return self.env.reset(**kw)

    def step(self, action):
        ob, rew, done, info = self.env.step(action)
        self.rewards.append(rew)
        if done:
            info = dict(info)
            info["episode"] = {"r": round(float(sum(self.rewards)), 6), "l": len(self.rewards), "t": 0.0}
        return ob, rew, done, info


class CloudpickleWrapper:
    def __init__(self, x):
        self.x = x


@contextmanager
def clear_mpi_env_vars():
    yield


class _Logger:
    @staticmethod
    def log(*a, **k):
        pass

    warn = log

    @staticmethod
    @contextmanager
    def scoped_configure(**k):
        yield


_INSTALLED = False


def install_shims():
    """Idempotently register every stand-in and put the reference on sys.path."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT} (it only exists in the build container)")
    if not hasattr(np, "bool"):  # declared oracle patch P2 (shmem_vec_env.py:22)
        np.bool = bool

    gym = _module("gym", Env=Env, Wrapper=Wrapper, ObservationWrapper=ObservationWrapper, make=_make)
    spaces = _module("gym.spaces", Box=Box, Dict=Dict)
    _module("gym.spaces.box", Box=Box)
    _module("gym.spaces.dict", Dict=Dict)
    envs_m = _module("gym.envs")
    _module("gym.envs.registration", register=_register)
    gym.spaces, gym.envs = spaces, envs_m

    _module("baselines", logger=_Logger)
    _module("baselines.logger", log=_Logger.log, warn=_Logger.warn, scoped_configure=_Logger.scoped_configure)
    _module("baselines.bench", Monitor=Monitor)
    _module("baselines.common")
    _module("baselines.common.atari_wrappers", make_atari=None, wrap_deepmind=None)
    _module("baselines.common.vec_env", VecEnvWrapper=VecEnvWrapper, VecEnv=VecEnv)
    _module("baselines.common.vec_env.dummy_vec_env", DummyVecEnv=DummyVecEnv)
    _module("baselines.common.vec_env.vec_normalize", VecNormalize=VecEnvWrapper)
    _module("baselines.common.vec_env.vec_env", VecEnv=VecEnv, CloudpickleWrapper=CloudpickleWrapper,
            clear_mpi_env_vars=clear_mpi_env_vars)
    _module("baselines.common.vec_env.util", dict_to_obs=dict_to_obs, obs_space_info=obs_space_info,
            obs_to_dict=obs_to_dict)

    for name in ("matplotlib", "matplotlib.legend", "matplotlib.lines", "matplotlib.pyplot", "matplotlib.text",
                 "matplotlib.cm", "matplotlib.patches", "matplotlib.animation", "matplotlib.colors"):
        m = _Inert(name)
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], child, m)

    sys.path.insert(0, os.path.join(_HERE, "refshim"))
    import shapely_standin

    shapely_standin.install()
    import rvo2  # noqa: F401  (oracle/refshim/rvo2.py)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _INSTALLED = True

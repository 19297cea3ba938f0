"""Minimal stand-ins for the two gym space types the reference's API exposes
(`gym.spaces.Box`, `gym.spaces.Dict`, crowd_sim/envs/crowd_sim_dict.py:31-69).
The real gym classes are used when gym is importable."""
import numpy as np

try:  # pragma: no cover - gym is not installed in the build image
    from gym.spaces import Box, Dict  # type: ignore
except Exception:  # noqa: BLE001

    class Box(object):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(np.shape(low) if shape is None else shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape)
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape)

        def __repr__(self):
            return "Box%s" % (self.shape,)

    class Dict(object):
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, key):
            return self.spaces[key]

        def __repr__(self):
            return "Dict(%s)" % ", ".join("%s: %r" % kv for kv in self.spaces.items())


def crowd_spaces(human_num):
    """observation / action spaces of CrowdSimDict.set_robot (crowd_sim_dict.py:24-69)."""
    inf = np.inf
    obs = Dict({
        "robot_node": Box(-inf, inf, shape=(1, 7), dtype=np.float32),
        "temporal_edges": Box(-inf, inf, shape=(1, 2), dtype=np.float32),
        "spatial_edges": Box(-inf, inf, shape=(human_num, 2), dtype=np.float32),
    })
    act = Box(-inf * np.ones(2), inf * np.ones(2), dtype=np.float32)
    return obs, act

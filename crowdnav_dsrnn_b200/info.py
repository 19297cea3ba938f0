"""Event value types of the step info dict (crowd_sim/envs/utils/info.py:1-38): same
class names, `__str__` texts and the `Danger.min_dist` attribute, so callers'
`isinstance(info["info"]["event"], ReachGoal)` checks (train.py:268-276,
evaluation.py:211-260) keep working."""
import sys

from . import abi


class Timeout(object):
    def __str__(self):
        return "Timeout"


class ReachGoal(object):
    def __str__(self):
        return "Reaching goal"


class Danger(object):
    def __init__(self, min_dist):
        self.min_dist = min_dist

    def __str__(self):
        return "Too close"


class Collision(object):
    def __str__(self):
        return "Collision"


class Nothing(object):
    def __str__(self):
        return ""


def _classes():
    """When the caller is the reference code base itself (its `crowd_sim.envs.utils.info` is loaded), events are instances
    of ITS classes, so the `isinstance` checks of the unmodified train.py / evaluation.py hold; otherwise the twins above."""
    ref = sys.modules.get("crowd_sim.envs.utils.info")
    if ref is not None and all(hasattr(ref, n) for n in ("Timeout", "ReachGoal", "Danger", "Collision", "Nothing")):
        return ref
    return sys.modules[__name__]


def make_event(code, dmin):
    m = _classes()
    if code == abi.EV_NOTHING:
        return m.Nothing()
    if code == abi.EV_DANGER:
        return m.Danger(dmin)
    if code == abi.EV_REACH_GOAL:
        return m.ReachGoal()
    if code == abi.EV_COLLISION:
        return m.Collision()
    if code == abi.EV_TIMEOUT:
        return m.Timeout()
    raise ValueError("unknown event code %r" % (code,))

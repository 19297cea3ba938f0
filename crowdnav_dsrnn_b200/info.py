"""Event value types of the step info dict (crowd_sim/envs/utils/info.py:1-38): same
class names, `__str__` texts and the `Danger.min_dist` attribute, so callers'
`isinstance(info["info"]["event"], ReachGoal)` checks (train.py:268-276,
evaluation.py:211-260) keep working."""
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


def make_event(code, dmin):
    if code == abi.EV_NOTHING:
        return Nothing()
    if code == abi.EV_DANGER:
        return Danger(dmin)
    if code == abi.EV_REACH_GOAL:
        return ReachGoal()
    if code == abi.EV_COLLISION:
        return Collision()
    if code == abi.EV_TIMEOUT:
        return Timeout()
    raise ValueError("unknown event code %r" % (code,))

"""TEST INFRASTRUCTURE -- stand-in for the third-party `rvo2` module.

The reference imports `rvo2` at crowd_nav/policy/orca.py:2 (Python-RVO2, a
Cython wrapper around RVO2 C++, un-pinned and not installable here).  This
module exposes the exact `PyRVOSimulator` method set used at
crowd_nav/policy/orca.py:87-136 on top of oracle/rvo2_port.c (restated RVO2;
parity UNPINNED against the real wheel).

Python floats are narrowed to C `float` at this boundary exactly as the Cython
wrapper does (`Vector2(float, float)`).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "..", "_build", "librvo2port.so")
if not os.path.exists(_LIB_PATH):
    raise ImportError(f"{_LIB_PATH} missing: run `make -C oracle`")
_lib = ctypes.CDLL(_LIB_PATH)

_f, _z, _p = ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p
_lib.rvo_create.restype = _p
_lib.rvo_create.argtypes = [_f]
_lib.rvo_destroy.argtypes = [_p]
_lib.rvo_add_agent.restype = ctypes.c_long
_lib.rvo_add_agent.argtypes = [_p, _f, _f, _f, _z, _f, _f, _f, _f, _f]
_lib.rvo_num_agents.restype = _z
_lib.rvo_num_agents.argtypes = [_p]
for _name in ("rvo_set_position", "rvo_set_velocity", "rvo_set_pref_velocity"):
    getattr(_lib, _name).argtypes = [_p, _z, _f, _f]
_lib.rvo_get_velocity.argtypes = [_p, _z, ctypes.POINTER(_f)]
_lib.rvo_get_position.argtypes = [_p, _z, ctypes.POINTER(_f)]
_lib.rvo_do_step.restype = ctypes.c_int
_lib.rvo_do_step.argtypes = [_p]


class PyRVOSimulator:
    def __init__(self, timeStep, neighborDist, maxNeighbors, timeHorizon, timeHorizonObst,
                 radius, maxSpeed, velocity=(0.0, 0.0)):
        self._h = _lib.rvo_create(timeStep)
        self._defaults = (neighborDist, maxNeighbors, timeHorizon, timeHorizonObst, radius, maxSpeed, velocity)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.rvo_destroy(h)

    def addAgent(self, pos, neighborDist=None, maxNeighbors=None, timeHorizon=None, timeHorizonObst=None,
                 radius=None, maxSpeed=None, velocity=None):
        d = self._defaults
        neighborDist = d[0] if neighborDist is None else neighborDist
        maxNeighbors = d[1] if maxNeighbors is None else maxNeighbors
        timeHorizon = d[2] if timeHorizon is None else timeHorizon
        radius = d[4] if radius is None else radius
        maxSpeed = d[5] if maxSpeed is None else maxSpeed
        velocity = d[6] if velocity is None else velocity
        return _lib.rvo_add_agent(self._h, pos[0], pos[1], neighborDist, int(maxNeighbors), timeHorizon,
                                  radius, maxSpeed, velocity[0], velocity[1])

    def getNumAgents(self):
        return _lib.rvo_num_agents(self._h)

    def setAgentPosition(self, i, pos):
        _lib.rvo_set_position(self._h, i, pos[0], pos[1])

    def setAgentVelocity(self, i, vel):
        _lib.rvo_set_velocity(self._h, i, vel[0], vel[1])

    def setAgentPrefVelocity(self, i, vel):
        _lib.rvo_set_pref_velocity(self._h, i, vel[0], vel[1])

    def doStep(self):
        if _lib.rvo_do_step(self._h) != 0:
            raise RuntimeError("rvo2 port: too many agents")

    def getAgentVelocity(self, i):
        out = (_f * 2)()
        _lib.rvo_get_velocity(self._h, i, out)
        return (float(out[0]), float(out[1]))

    def getAgentPosition(self, i):
        out = (_f * 2)()
        _lib.rvo_get_position(self._h, i, out)
        return (float(out[0]), float(out[1]))

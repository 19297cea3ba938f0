"""TEST INFRASTRUCTURE -- drives the reference's OWN CrowdSimDict (imported from
/root/reference under oracle/ref_import.py shims) on injected states.  Exists
only in the build container; used by oracle/gen_golden.py to produce
tests/golden/*.npz and by tests that are skipped when the reference is absent.

Declared oracle patches:
  P1  unicycle: crowd_sim.py:1004-1005,1023 read `.vx/.vy` from an ActionRot and
      crash.  The harness hands the env an ActionRot subclass instance carrying
      (vx, vy) = (v cos(theta+r), v sin(theta+r)), the velocity Agent.step will
      assign (agent.py:206-212).  Only SM4/SM5 info fields read them.
  P3  state injection through the reference's own setters (Agent.set_list) and
      attributes; fresh Human objects so each ORCA policy rebuilds its simulator.
No reference source line is modified.
"""
import math

import numpy as np

from . import ref_import

SCENARIOS = ["circle_crossing", "square_crossing", "parallel_traffic", "perpendicular_traffic",
             "side_pref_passing", "side_pref_overtaking", "side_pref_crossing"]
EVENT_CODE = {"Nothing": 0, "Danger": 1, "ReachGoal": 2, "Collision": 3, "Timeout": 4}


def make_reference_config(**over):
    """Fresh instance-level copy of the reference Config with dotted overrides, e.g. {'sim.human_num': 10}."""
    ref_import.install_shims()
    import copy

    from crowd_nav.configs.config import Config

    cfg = Config()
    # the reference keeps sections as CLASS attributes: deep-copy them onto the instance
    for name in dir(Config):
        if name.startswith("_"):
            continue
        setattr(cfg, name, copy.deepcopy(getattr(Config, name)))
    for key, val in over.items():
        sec, _, attr = key.partition(".")
        setattr(getattr(cfg, sec), attr, val)
    return cfg


class RefEnv:
    """One reference CrowdSimDict with injection / extraction helpers."""

    def __init__(self, config, n_envs=2, rank=0, phase=None):
        ref_import.install_shims()
        import crowd_sim  # noqa: F401  (registers the gym ids)
        from crowd_sim.envs import crowd_sim_dict as csd
        from crowd_sim.envs.utils.action import ActionRot

        self._csd = csd
        self._ActionRot = ActionRot
        self.config = config
        self.env = csd.CrowdSimDict()
        self.env.configure(config)
        self.env.thisSeed = config.env.seed + rank
        self.env.nenv = n_envs
        self.env.phase = phase or ("train" if n_envs > 1 else "test")
        self.H = config.sim.human_num

        env = self.env

        class _ActionRotP1(ActionRot):
            pass

        def make_action_rot(v, r):  # declared patch P1
            a = _ActionRotP1(v, r)
            a.vx = v * np.cos(env.robot.theta + r)
            a.vy = v * np.sin(env.robot.theta + r)
            return a

        self._make_action_rot = make_action_rot

    # ------------------------------------------------------------------ injection (P3)
    def inject(self, robot9, humans9, belief5, extras4, counters4):
        from crowd_sim.envs.utils.human import Human

        env, cfg = self.env, self.config
        f = lambda x: np.float64(x)
        env.robot.set_list(*[f(v) for v in robot9])
        env.humans = []
        for row in humans9:
            h = Human(cfg, "humans")
            h.set_list(*[f(v) for v in row])
            env.humans.append(h)
        env.last_human_states = np.array(belief5, dtype=np.float64).reshape(self.H, 5).copy()
        env.desiredVelocity = [f(extras4[0]), 0.0]
        env.potential = f(extras4[1])
        env.last_acceleration = (f(extras4[2]), f(extras4[3]))
        t = 0
        for _ in range(int(counters4[0])):
            t += env.time_step
        env.global_time = t
        env.scenario_counter = int(counters4[1])
        env.case_counter[env.phase] = int(counters4[2])
        env.current_scenario = SCENARIOS[int(counters4[3])]

    def extract(self):
        env = self.env
        robot = np.array(env.robot.get_full_state_list(), dtype=np.float64)
        humans = np.array([h.get_full_state_list() for h in env.humans], dtype=np.float64)
        belief = np.array(env.last_human_states, dtype=np.float64)
        extras = np.array([env.desiredVelocity[0], env.potential, env.last_acceleration[0],
                           env.last_acceleration[1]], dtype=np.float64)
        return dict(robot=robot, humans=humans, belief=belief, extras=extras, global_time=float(env.global_time))

    # ------------------------------------------------------------------ stepping
    def step(self, action2):
        """Reference step on a float32 action (as VecPyTorch delivers it, envs.py:224-229)."""
        env = self.env
        action = np.array(action2, dtype=np.float32)
        if env.robot.kinematics == "unicycle":
            self._csd.ActionRot = self._make_action_rot
        try:
            ob, reward, done, info = env.step(action)
        finally:
            self._csd.ActionRot = self._ActionRot
        _, _, vis = env.get_num_human_in_fov()
        si = info["info"]
        ev = type(si["event"]).__name__
        out = dict(
            robot_node=np.asarray(ob["robot_node"], dtype=np.float64),
            temporal_edges=np.asarray(ob["temporal_edges"], dtype=np.float64),
            spatial_edges=np.asarray(ob["spatial_edges"], dtype=np.float64),
            visible=np.array(vis, dtype=bool),
            reward=float(reward), done=bool(done), event=EVENT_CODE[ev],
            dmin=float(si["event"].min_dist) if ev == "Danger" else math.nan,
            aggregate_nav_time=float(si["aggregate_nav_time"]), path_violation=float(si["path_violation"]),
            personal_violation=float(si["personal_violation"]), jerk_cost=float(si["jerk_cost"]),
            dist_to_goal=float(si["dist_to_goal"]), speed_violation=float(si["speed_violation"]),
            scenario=SCENARIOS.index(si["scenario"]),
        )
        if self.config.test.side_preference:
            sp = si[si["scenario"]]
            out["side_left"], out["side_right"] = float(sp["left"]), float(sp["right"])
            out["separation"] = float(si["separation"])
        return out

    def reset(self):
        ob = self.env.reset()
        return {k: np.asarray(v, dtype=np.float64) for k, v in ob.items()}

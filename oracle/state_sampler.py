"""TEST INFRASTRUCTURE -- seeded generator of injected env states / actions that
exercise the branches of the crowd step (collision, danger band, goal reached,
timeout, wall contact, ORCA overlap / infeasible LPs, >10 m neighbour cut-off,
limited FOV, zero velocities, tiny unicycle rotations).  Everything is rounded
to float32 so the reference (float64 Python), the C oracle and the CUDA kernel
start from bit-identical inputs.
"""
import numpy as np

from crowdnav_dsrnn_b200 import abi


def sample(cfg, n_envs, seed):
    """Returns dict(robot[N,9], humans[N,H,9], belief[N,H,5], extras[N,4], counters[N,4], action[N,2])."""
    rng = np.random.default_rng(seed)
    N, H = n_envs, cfg.human_num
    uni = cfg.kinematics == abi.UNICYCLE
    robot = np.zeros((N, 9))
    humans = np.zeros((N, H, 9))
    belief = np.zeros((N, H, 5))
    extras = np.zeros((N, 4))
    counters = np.zeros((N, 4), np.int64)
    action = np.zeros((N, 2))
    R = cfg.circle_radius
    for e in range(N):
        kind = e % 16
        # ---------------- robot
        p = rng.uniform(-R, R, 2)
        g = rng.uniform(-R, R, 2)
        if kind == 1:   # already at the goal
            g = p + rng.uniform(-0.15, 0.15, 2)
        if kind == 2:   # touching / near a wall
            side = rng.integers(0, 4)
            off = cfg.square_width / 2 - cfg.robot_radius + rng.uniform(-0.05, 0.05)
            p = rng.uniform(-8, 8, 2)
            p[side % 2] = off if side < 2 else -off
        sp = rng.uniform(0, 1.0) if kind != 3 else 0.0
        ang = rng.uniform(0, 2 * np.pi)
        theta = rng.uniform(0, 2 * np.pi) if uni else np.pi / 2
        v = sp * np.array([np.cos(theta), np.sin(theta)]) if uni else sp * np.array([np.cos(ang), np.sin(ang)])
        robot[e] = [p[0], p[1], v[0], v[1], cfg.robot_radius, g[0], g[1], cfg.robot_v_pref, theta]
        # ---------------- humans
        for i in range(H):
            r = rng.uniform(0.3, 0.5)
            vp = rng.uniform(0.5, 1.5)
            mode = rng.uniform()
            if mode < 0.55:
                hp = rng.uniform(-R - 1, R + 1, 2)
            elif mode < 0.75:  # around the robot: collisions and the danger band
                d = cfg.robot_radius + r + rng.uniform(-0.08, 0.45)
                a = rng.uniform(0, 2 * np.pi)
                hp = p + d * np.array([np.cos(a), np.sin(a)])
            elif mode < 0.9 and i > 0:  # around an earlier human: ORCA overlap / tight LPs
                k = rng.integers(0, i)
                d = humans[e, k, 4] + r + 0.32 + rng.uniform(-0.25, 0.6)
                a = rng.uniform(0, 2 * np.pi)
                hp = humans[e, k, 0:2] + d * np.array([np.cos(a), np.sin(a)])
            else:  # far away: beyond ORCA's 10 m neighbour range
                hp = rng.uniform(-22, 22, 2)
            hv = rng.uniform(0, vp) * np.array([np.cos(a2 := rng.uniform(0, 2 * np.pi)), np.sin(a2)])
            if rng.uniform() < 0.1:
                hv = np.zeros(2)
            hg = rng.uniform(-R - 1, R + 1, 2)
            if rng.uniform() < 0.1:   # standing on its goal
                hg = hp + rng.uniform(-0.2, 0.2, 2)
            humans[e, i] = [hp[0], hp[1], hv[0], hv[1], r, hg[0], hg[1], vp, 0.0]
            belief[e, i] = [hp[0] + rng.normal(0, 0.3), hp[1] + rng.normal(0, 0.3), hv[0], hv[1], r]
        # ---------------- extras / counters / action
        dgoal = np.linalg.norm(p - g)
        extras[e] = [rng.uniform(-1, 1) if uni else 0.0, -dgoal + rng.normal(0, 0.05),
                     rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5)]
        steps = rng.integers(0, max(cfg.timeout_step, 1))
        if kind == 4:
            steps = cfg.timeout_step
        if kind == 5:
            steps = cfg.timeout_step - 1
        if kind == 6:   # a step after which global_time % 5 == 0 may fire
            steps = 19
        scen = cfg.scenarios[rng.integers(0, cfg.n_scenarios)]
        counters[e] = [steps, rng.integers(1, 1000), rng.integers(0, 100000), scen]
        if uni:
            action[e] = rng.normal(0, 0.08, 2)
            if kind == 7:
                action[e, 1] = rng.uniform(-1e-4, 1e-4)   # |r| around the 1e-4 "no translation" epsilon
        else:
            action[e] = rng.normal(0, 0.7, 2)
            if kind == 8:
                action[e] = 0.0
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    return dict(robot=f32(robot), humans=f32(humans), belief=f32(belief), extras=f32(extras),
                counters=np.ascontiguousarray(counters, np.int32), action=f32(action))

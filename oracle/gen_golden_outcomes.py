"""TEST INFRASTRUCTURE -- episode-level outcome statistics of the REFERENCE (its own CrowdSimDict + its own Policy,
shipped checkpoints, deterministic actions) for the distributional parity check of BASELINE.json's north_star:
"over 2000-episode test runs, success/collision/timeout rates must fall within binomial noise of the reference".

Build container only (imports /root/reference under oracle/ref_import.py shims; rvo2 = restated RVO2):

    python -m oracle.gen_golden_outcomes            # writes tests/golden/outcomes_<case>.json

Each worker process plays `episodes / workers` test-phase episodes the way evaluation.py:96-191 does (1 env,
`phase="test"`, hidden state carried over, mask 0 after a terminal step), with `env.seed` offset per worker so the
workers' case seeds (1000 + case_counter + seed) do not overlap.
"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

CASES = {
    # the optional human behaviours -- social-force humans; human 0 randomly blind with radius / v_pref
    # drifting at every new end goal.  (humans.random_policy_changing cannot run in the reference: crowd_sim.py:472-473 calls
    # the policy constructors without their config and a Human.set_policy that does not exist.)
    "n4_social_force_h5": dict(over={"humans.policy": "social_force"}, ckpt="data/example_model/checkpoints/27776.pt", episodes=2000),
    # the group environment: circles of static humans plus up to 4 walkers (crowd_sim.py:476-622)
    "n4_group_h8": dict(over={"sim.group_human": True, "sim.human_num": 8}, ckpt="data/example_model/checkpoints/27776.pt", episodes=2000),
    "n4_options_h5": dict(over={"humans.random_unobservability": True, "humans.unobservable_chance": 0.5,
                                "humans.random_radii": True, "humans.random_v_pref": True},
                          ckpt="data/example_model/checkpoints/27776.pt", episodes=2000),
    # BASELINE.json configs[3]: test.social_metrics=True (circle_radius 4, sequential scenarios), holonomic, 5 humans
    "social_h5_holonomic": dict(over={"test.social_metrics": True, "sim.circle_radius": 4, "env.test_size": 2000},
                                ckpt="data/example_model/checkpoints/27776.pt", episodes=2000),
    # reference defaults (circle_radius 6, random scenario per episode), holonomic, 5 humans
    "default_h5_holonomic": dict(over={}, ckpt="data/example_model/checkpoints/27776.pt", episodes=2000),
    # unicycle checkpoint with its own time step (data/example_model_unicycle/configs/config.py:20); needs oracle patch P1
    "default_h5_unicycle": dict(over={"action_space.kinematics": "unicycle", "env.time_step": 0.1,
                                      "reward.discomfort_penalty_factor": 1.0},
                                ckpt="data/example_model_unicycle/checkpoints/55554.pt", episodes=2000),
    # BASELINE.json configs[1]: unicycle robot, 10 humans, the four mixed scenarios, dt 0.1 (the reference still terminates
    # at H=10: the unbounded goal re-sampling loop of crowd_sim.py:787-806 only becomes prohibitive near H=20)
    "c2_h10_unicycle": dict(over={"action_space.kinematics": "unicycle", "env.time_step": 0.1, "sim.human_num": 10,
                                  "reward.discomfort_penalty_factor": 1.0,
                                  "sim.test_sim": ["circle_crossing", "square_crossing", "parallel_traffic", "perpendicular_traffic"]},
                            ckpt="data/example_model_unicycle/checkpoints/55554.pt", episodes=1200),
    # BASELINE.json configs[3], second half: side-preference scenarios (1 human, circle_radius 4, 200 episodes each;
    # 2000 here to tighten the noise), config.py:27-40,51-54,87-92
    "sidepref_passing": dict(over={"test.side_preference": True, "sim.test_sim": ["side_pref_passing"], "sim.circle_radius": 4,
                                   "sim.human_num": 1, "env.test_size": 200, "humans.random_goal_changing": False,
                                   "humans.end_goal_changing": False},
                             ckpt="data/example_model/checkpoints/27776.pt", episodes=2000),
    "sidepref_overtaking": dict(over={"test.side_preference": True, "sim.test_sim": ["side_pref_overtaking"], "sim.circle_radius": 4,
                                      "sim.human_num": 1, "env.test_size": 200, "humans.random_goal_changing": False,
                                      "humans.end_goal_changing": False},
                                ckpt="data/example_model/checkpoints/27776.pt", episodes=2000),
    "sidepref_crossing": dict(over={"test.side_preference": True, "sim.test_sim": ["side_pref_crossing"], "sim.circle_radius": 4,
                                    "sim.human_num": 1, "env.test_size": 200, "humans.random_goal_changing": False,
                                    "humans.end_goal_changing": False},
                              ckpt="data/example_model/checkpoints/27776.pt", episodes=2000),
}


def _worker(args):
    name, rank, n_workers, episodes = args
    import torch

    from . import ref_harness, ref_import

    torch.set_num_threads(1)
    ref_import.install_shims()
    from pytorchBaselines.a2c_ppo_acktr.model import Policy

    case = CASES[name]
    over = dict(case["over"])
    over["training.cuda"] = False
    over["training.num_processes"] = 1
    over["env.seed"] = 100000 * rank        # disjoint case seeds per worker
    cfg = ref_harness.make_reference_config(**over)
    renv = ref_harness.RefEnv(cfg, n_envs=1, phase="test")
    env = renv.env
    env.scenario_counter = (episodes * rank) % 4 if cfg.test.social_metrics else 0
    H = cfg.sim.human_num
    unicycle = cfg.action_space.kinematics == "unicycle"
    spaces = {"robot_node": ref_import.Box(-np.inf, np.inf, (1, 7)), "temporal_edges": ref_import.Box(-np.inf, np.inf, (1, 2)),
              "spatial_edges": ref_import.Box(-np.inf, np.inf, (H, 2))}
    policy = Policy(spaces, ref_import.Box(-np.inf, np.inf, (2,)), base="srnn", base_kwargs=cfg)
    policy.load_state_dict(torch.load(os.path.join(ref_import.REFERENCE_ROOT, case["ckpt"]), map_location="cpu"))
    policy.eval()
    hx = {"human_node_rnn": torch.zeros(1, 1, 128), "human_human_edge_rnn": torch.zeros(1, H + 1, 256)}
    mask = torch.zeros(1, 1)
    t = lambda a, shape: torch.as_tensor(np.asarray(a, dtype=np.float32)).reshape(shape)
    rows = []
    ob = env.reset()
    for _ in range(episodes):
        done, steps, ret = False, 0, 0.0
        sm = dict(personal_violation=0.0, path_violation=0.0, aggregate_nav_time=0.0, jerk_cost=0.0, speed_violation=0.0,
                  left=0.0, right=0.0)
        while not done:
            obs = {"robot_node": t(ob["robot_node"], (1, 1, 7)), "temporal_edges": t(ob["temporal_edges"], (1, 1, 2)),
                   "spatial_edges": t(ob["spatial_edges"], (1, H, 2))}
            with torch.no_grad():
                _, action, _, hx = policy.act(obs, hx, mask, deterministic=True)
            if unicycle:                      # declared oracle patch P1 (see oracle/ref_harness.py)
                renv._csd.ActionRot = renv._make_action_rot
            try:
                ob, reward, done, info = env.step(action[0].numpy().copy())
            finally:
                renv._csd.ActionRot = renv._ActionRot
            steps += 1
            ret += float(reward)
            si = info["info"]                 # per-step accounting of evaluation.py:155-190 (counts, not yet times dt)
            for k in ("personal_violation", "path_violation", "aggregate_nav_time", "jerk_cost", "speed_violation"):
                sm[k] += float(si[k])
            if cfg.test.side_preference:
                sm["left"] += si[si["scenario"]]["left"]
                sm["right"] += si[si["scenario"]]["right"]
            mask = torch.tensor([[0.0 if done else 1.0]])
        ev = type(info["info"]["event"]).__name__
        rows.append((ev, info["info"]["scenario"], steps, ret, sm))
        ob = env.reset()
    return rows


def run_case(name, workers):
    case = CASES[name]
    per = case["episodes"] // workers
    t0 = time.time()
    with mp.get_context("fork").Pool(workers) as pool:
        parts = pool.map(_worker, [(name, r, workers, per) for r in range(workers)])
    rows = [r for p in parts for r in p]
    n = len(rows)
    ev = [r[0] for r in rows]
    stats = {
        "episodes": n,
        "success": ev.count("ReachGoal") / n, "collision": ev.count("Collision") / n, "timeout": ev.count("Timeout") / n,
        "mean_steps": float(np.mean([r[2] for r in rows])),
        "mean_steps_success": float(np.mean([r[2] for r in rows if r[0] == "ReachGoal"])),
        "mean_return": float(np.mean([r[3] for r in rows])),
        "per_scenario": {},
        "social": {k: {"mean": float(np.mean([r[4][k] for r in rows])), "std": float(np.std([r[4][k] for r in rows]))}
                   for k in ("personal_violation", "path_violation", "aggregate_nav_time", "jerk_cost", "speed_violation")},
        "side_left_episodes": sum(r[4]["left"] > r[4]["right"] for r in rows) / n,
        "side_right_episodes": sum(r[4]["right"] > r[4]["left"] for r in rows) / n,
        "overrides": case["over"], "checkpoint": case["ckpt"], "wall_seconds": time.time() - t0,
    }
    for scn in sorted(set(r[1] for r in rows)):
        sub = [r for r in rows if r[1] == scn]
        stats["per_scenario"][scn] = {"episodes": len(sub), "success": sum(r[0] == "ReachGoal" for r in sub) / len(sub),
                                      "collision": sum(r[0] == "Collision" for r in sub) / len(sub),
                                      "timeout": sum(r[0] == "Timeout" for r in sub) / len(sub)}
    with open(os.path.join(GOLDEN_DIR, "outcomes_%s.json" % name), "w") as f:
        json.dump(stats, f, indent=1)
    print(name, {k: stats[k] for k in ("episodes", "success", "collision", "timeout", "mean_steps", "mean_return", "wall_seconds")})


if __name__ == "__main__":
    from . import crowd_oracle

    crowd_oracle.build()
    for name in (sys.argv[1:] or list(CASES)):
        run_case(name, workers=8)

"""Batched evaluator: the test loop of pytorchBaselines/evaluation.py:96-330 with the
`env.test_size` episodes played as ONE batch on the GPU instead of one after the other in a single env.

Env k of the batch plays what would have been the k-th test episode of the reference's single env: its
`scenario_counter` starts at k (sequential scenarios under test.social_metrics, crowd_sim_dict.py:114-122) and its
`case_counter` at k (seed = 1000 + k + env.seed, crowd_sim_dict.py:147-154).  The reference carries
`last_acceleration` from one episode into the next (it is initialised in `configure` only, crowd_sim.py:208), which
shows up in the first jerk sample (SM4) of every episode; to reproduce that, every env first plays `warmup_episodes`
unscored episodes (env k starts at counter k - warmup) and the NEXT one is scored.  The rollout stops when every env
has finished its scored episode.  Per-step social metrics (SM1-SM6, evaluation.py:155-190) are accumulated on the
device from the step-info columns.
"""
import torch

from . import abi
from .envs import CrowdVecEnv


@torch.no_grad()
def evaluate_batched(actor_critic, config, device, episodes=None, seed=None, deterministic=True, warmup_episodes=1):
    n = int(config.env.test_size if episodes is None else episodes)
    device = torch.device(device)
    venv = CrowdVecEnv(config, n, device, seed=config.env.seed if seed is None else seed, phase="test", nenv=1)
    eng = venv.engine
    H = config.sim.human_num
    st = eng.get_state()
    W = int(warmup_episodes)
    size = max(1, int(eng.cfg.case_size))
    ks = torch.arange(n, dtype=torch.int64, device=device) - W
    st["counters"][:, 1] = (ks % 4 + 4).to(torch.int32)          # same residue mod 4, never negative
    st["counters"][:, 2] = (ks % size).to(torch.int32)
    eng.set_state(counters=st["counters"])
    obs = eng.reset().obs()
    hx = {"human_node_rnn": torch.zeros(n, 1, 128, device=device),
          "human_human_edge_rnn": torch.zeros(n, H + 1, 256, device=device)}
    masks = torch.zeros(n, 1, device=device)
    finished = torch.zeros(n, dtype=torch.bool, device=device)
    ep_idx = torch.zeros(n, dtype=torch.int32, device=device)      # episodes completed so far by each env
    event = torch.zeros(n, dtype=torch.int32, device=device)
    scenario = torch.zeros(n, dtype=torch.int32, device=device)
    length = torch.zeros(n, dtype=torch.int32, device=device)
    ret = torch.zeros(n, device=device)
    col = abi.INFO_COLUMNS
    acc = {k: torch.zeros(n, device=device) for k in ("personal_violation", "path_violation", "aggregate_nav_time",
                                                       "jerk_cost", "speed_violation", "side_left", "side_right")}
    path_len = torch.zeros(n, device=device)
    last_pos = obs["robot_node"][:, 0, 0:2].clone()
    for _ in range((W + 1) * (int(eng.cfg.timeout_step) + 2)):
        _, action, _, hx = actor_critic.act(obs, hx, masks, deterministic=deterministic)
        obs, _, done, buf = venv.step_device(action)
        live = ep_idx == W                                          # currently playing the scored episode
        for k in acc:
            acc[k] += torch.where(live, buf.info[:, col[k]], torch.zeros_like(ret))
        pos = obs["robot_node"][:, 0, 0:2]
        path_len += torch.where(live, (pos - last_pos).norm(dim=-1), torch.zeros_like(ret))   # evaluation.py:134-143
        last_pos = pos.clone()
        first = live & done.bool()
        ep_idx += done.to(torch.int32)
        event = torch.where(first, buf.event, event)
        scenario = torch.where(first, buf.scenario, scenario)
        length = torch.where(first, buf.episode_length, length)
        ret = torch.where(first, buf.episode_return, ret)
        finished |= first
        masks = buf.not_done                                        # 1 - done, written by the step kernel
        if bool(finished.all()):
            break
    venv.close()
    ev = event.cpu()
    scn = scenario.cpu()
    steps = length.cpu().float()
    dt = float(config.env.time_step)
    out = {
        "episodes": n,
        "success": float((ev == abi.EV_REACH_GOAL).float().mean()),
        "collision": float((ev == abi.EV_COLLISION).float().mean()),
        "timeout": float((ev == abi.EV_TIMEOUT).float().mean()),
        "unfinished": int((~finished).sum()),
        "mean_steps": float(steps.mean()),
        "mean_steps_success": float(steps[ev == abi.EV_REACH_GOAL].mean()) if bool((ev == abi.EV_REACH_GOAL).any()) else float("nan"),
        "mean_nav_time_success": float(steps[ev == abi.EV_REACH_GOAL].mean()) * dt if bool((ev == abi.EV_REACH_GOAL).any()) else float("nan"),
        "mean_return": float(ret.mean()),
        "mean_path_length": float(path_len.mean()),
        "per_scenario": {},
    }
    for k, v in acc.items():                     # per-episode sums of the per-step values (evaluation.py:155-190 multiplies
        out["mean_" + k + "_per_episode"] = float(v.mean())   # the violation counts by time_step when it logs them)
        out["std_" + k + "_per_episode"] = float(v.std())
    # side preference: an episode is labelled with the side seen most often (evaluation.py:199-210)
    out["side_left_episodes"] = float((acc["side_left"] > acc["side_right"]).float().mean())
    out["side_right_episodes"] = float((acc["side_right"] > acc["side_left"]).float().mean())
    for s in sorted(set(scn.tolist())):
        sel = scn == s
        out["per_scenario"][abi.SCENARIOS[s]] = {
            "episodes": int(sel.sum()), "success": float((ev[sel] == abi.EV_REACH_GOAL).float().mean()),
            "collision": float((ev[sel] == abi.EV_COLLISION).float().mean()),
            "timeout": float((ev[sel] == abi.EV_TIMEOUT).float().mean())}
    return out

"""Distributional parity of the reset (SURVEY.md test plan T5; DESIGN.md deviations D1 RNG, D2 bounded rejection): spawn
positions, goals, attributes, min-distance constraints and scenario frequencies of the C oracle's reset (CPU) and of
`crowd_reset_kernel` (GPU) against the empirical distributions of the reference's OWN reset
(crowd_sim_dict.py:105-203 -> generate_robot_humans, crowd_sim.py:555-663), 3000 resets per case executed in the build
container by oracle/gen_golden_spawn.py -> tests/golden/spawn_*.npz (257 quantiles per feature).

Kolmogorov-Smirnov distance of every feature below the two-sample critical value at alpha = 1e-3 (humans of one env are
not independent, hence the factor 1.5 on per-human features); the kernel and the oracle are bit-identical to each other
(tests/test_gpu_crowd_step.py), so the CPU test already pins the algorithm and the GPU test pins the kernel."""
import glob
import math
import os

import numpy as np
import pytest

from crowdnav_dsrnn_b200 import abi
from helpers import GOLDEN, config_from_overrides
from oracle import crowd_oracle
from oracle.gen_golden_spawn import features

CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLDEN, "spawn_*.npz")))
N_ENVS = 4096


def _config(d):
    over = {k: eval(v) for k, v in zip(d["overrides_keys"].tolist(), d["overrides_vals"].tolist())}   # reprs written by the generator
    return config_from_overrides(over)


def _check(d, robot, humans, scenario):
    grid = d["quantile_grid"]
    samples = {}
    for e in range(robot.shape[0]):
        for k, v in features(robot[e].astype(np.float64), humans[e].astype(np.float64)).items():
            samples.setdefault(k, []).append(v)
    report = {}
    for k, parts in samples.items():
        x = np.sort(np.concatenate(parts))
        q_ref, n_ref = d["q_" + k], int(d["n_" + k])
        if q_ref[-1] - q_ref[0] < 1e-9:                       # constant in the reference (e.g. radius without randomisation)
            assert np.abs(x - q_ref[0]).max() < 1e-5, k
            continue
        # our CDF at the reference's quantiles (+1e-5: our state is float32, an atom such as radius 0.3 or the 11 m robot-goal distance sits an ulp or two off)
        cdf = np.searchsorted(x, q_ref + 1e-5, side="right") / x.size
        # an atom of the distribution (the static humans of the group environment: v_pref = 0, goal = position) repeats one
        # value over a run of quantiles: the reference CDF at that value is the END of the run
        last_of_run = np.r_[np.abs(q_ref[1:] - q_ref[:-1]) > 1e-5, True]
        idx = np.arange(grid.size)
        end = np.where(last_of_run, idx, grid.size)          # index of the run's end, propagated backwards
        end = np.minimum.accumulate(end[::-1])[::-1]
        dist = float(np.abs(cdf - grid[end])[1:-1].max())
        crit = 1.95 * math.sqrt((x.size + n_ref) / (x.size * n_ref)) + 1.0 / (grid.size - 1)
        if k.startswith("human_"):
            crit *= 1.5
        report[k] = (dist, crit)
        assert dist <= crit, (k, dist, crit)
    # rejection rule of the spawn (crowd_sim.py:378-390, 630-651): no pair closer than the discomfort distance, except for
    # the (rare) envs whose bounded rejection loop ran out of tries (deviation D2)
    for k in ("min_clear_human_human", "min_clear_robot_human"):
        if k in samples and d["q_" + k].size:
            x = np.concatenate(samples[k])
            assert (x < d["q_" + k][0] - 1e-3).mean() <= 2e-3, (k, float(x.min()), float(d["q_" + k][0]))
    # scenario choice: uniform over the configured scenarios, like random.choices (crowd_sim_dict.py:112-125)
    ref_counts = d["scenario_counts"].astype(np.float64)
    used = ref_counts > 0
    got = np.bincount(scenario, minlength=4).astype(np.float64)
    assert (got[~used] == 0).all()
    p = 1.0 / used.sum()
    for c in got[used]:
        assert abs(c / scenario.size - p) <= 4.0 * math.sqrt(p * (1 - p) / scenario.size) + 1e-12
    return report


@pytest.mark.parametrize("case", CASES)
def test_oracle_reset_matches_reference_spawn_distributions(case):
    d = np.load(os.path.join(GOLDEN, "spawn_%s.npz" % case))
    cfg_obj = _config(d)
    cfg = abi.flatten_config(cfg_obj, N_ENVS, phase="train")
    st = crowd_oracle.OracleState(N_ENVS, cfg.human_num)
    crowd_oracle.reset(cfg, st, n_threads=4)
    _check(d, st.robot, st.humans, st.counters[:, 3])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_reset_matches_reference_spawn_distributions(case):
    import torch

    from crowdnav_dsrnn_b200.engine import CrowdEngine

    d = np.load(os.path.join(GOLDEN, "spawn_%s.npz" % case))
    eng = CrowdEngine(_config(d), N_ENVS, torch.device("cuda:0"), phase="train")
    eng.reset()
    st = eng.get_state()
    _check(d, st["robot"].cpu().numpy(), st["humans"].cpu().numpy(), st["counters"][:, 3].cpu().numpy())
    eng.close()

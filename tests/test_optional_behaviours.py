"""CPU: the optional human behaviours of SURVEY 8(f) N4 in the C oracle -- social-force humans are pinned on golden vectors
of the reference's own SOCIAL_FORCE.predict (tests/golden/step_n4_*.npz, via test_oracle_golden.py); here the RNG-driven
switches (random_policy_changing, random_unobservability, random_radii / random_v_pref), which the reference draws from
its global MT19937 / `random` streams, are checked through the properties the reference's code implies."""
import numpy as np
import pytest

from crowdnav_dsrnn_b200 import Config, abi
from oracle import crowd_oracle, state_sampler


def _run(over, n=1024, H=6, seed=3, auto_reset=False):
    cfg_obj = Config(human_num=H)
    for k, v in over.items():
        sec, _, attr = k.partition(".")
        setattr(getattr(cfg_obj, sec), attr, v)
    cfg = abi.flatten_config(cfg_obj, n, phase="train")
    inp = state_sampler.sample(cfg, n, seed)
    act = inp.pop("action")
    st = crowd_oracle.OracleState(n, H)
    for f in ("robot", "humans", "belief", "extras", "counters"):
        getattr(st, f)[...] = inp[f]
    out = crowd_oracle.step(cfg, st, act, auto_reset=auto_reset, n_threads=4)
    return inp, st, out


def test_flags_reach_the_flat_config():
    c = Config(human_num=5)
    c.humans.policy = "social_force"
    c.humans.random_unobservability = True
    c.humans.unobservable_chance = 0.4
    c.sf.A, c.sf.B, c.sf.KI = 1.5, 0.8, 2.0
    f = abi.flatten_config(c, 4)
    assert f.human_policy == abi.POLICY_SOCIAL_FORCE and f.random_unobservability == 1 and f.unobservable_chance == 0.4
    assert (f.sf_A, f.sf_B, f.sf_KI) == (1.5, 0.8, 2.0)
    c.humans.policy = "cadrl"
    with pytest.raises(NotImplementedError):
        abi.flatten_config(c, 4)


def test_mixed_policies_are_a_per_episode_coin_flip():
    """random_policy_changing (crowd_sim.py:463-473): each human is ORCA or social force, equal chance; a human's velocity
    then equals one of the two pure-policy results."""
    _, st_orca, _ = _run({})
    _, st_sf, _ = _run({"humans.policy": "social_force"})
    _, st_mix, _ = _run({"humans.random_policy_changing": True})
    v_o, v_s, v_m = st_orca.humans[:, :, 2:4], st_sf.humans[:, :, 2:4], st_mix.humans[:, :, 2:4]
    is_o = np.all(v_m == v_o, axis=-1)
    is_s = np.all(v_m == v_s, axis=-1)
    assert np.all(is_o | is_s)
    distinct = ~np.all(v_o == v_s, axis=-1)
    frac_sf = (is_s & distinct).sum() / distinct.sum()
    assert abs(frac_sf - 0.5) < 4 * 0.5 / np.sqrt(distinct.sum())


def test_unobservability_only_touches_human_zero():
    _, st_ref, _ = _run({})
    _, st_un, _ = _run({"humans.random_unobservability": True, "humans.unobservable_chance": 1.0})
    # humans 1.. see everybody as before (their new GOALS may differ: the goal search avoids human 0's new position)
    assert np.array_equal(st_ref.humans[:, 1:, 0:4], st_un.humans[:, 1:, 0:4])
    changed = np.any(st_ref.humans[:, 0, 2:4] != st_un.humans[:, 0, 2:4], axis=-1)
    assert changed.mean() > 0.3                                               # human 0 reacts to dummies parked at (7, 7)
    _, st_zero, _ = _run({"humans.random_unobservability": True, "humans.unobservable_chance": -1.0})
    assert np.array_equal(st_ref.humans, st_zero.humans)                      # chance below every draw: nothing changes


def test_random_radii_and_v_pref_move_with_the_end_goal():
    """crowd_sim.py:779-786: radius / v_pref += U(-0.1, 0.1) exactly when a human is handed a new end goal."""
    inp, st, out = _run({"humans.random_radii": True, "humans.random_v_pref": True, "humans.random_goal_changing": False}, n=2048, H=5)
    d_r = st.humans[:, :, 4] - inp["humans"][:, :, 4]
    d_v = st.humans[:, :, 7] - inp["humans"][:, :, 7]
    reached = np.linalg.norm(inp["humans"][:, :, 5:7].astype(np.float64) - st.humans[:, :, 0:2], axis=-1) < inp["humans"][:, :, 4]
    moved = (d_r != 0) | (d_v != 0)
    assert moved.sum() > 200
    assert not np.any(moved & ~reached)                       # only humans standing on their (old) goal
    assert np.abs(d_r).max() <= 0.1 + 1e-6 and np.abs(d_v).max() <= 0.1 + 1e-6
    assert abs(float(d_r[moved].mean())) < 0.02 and abs(float(d_v[moved].mean())) < 0.02      # centred
    assert np.std(d_r[moved]) == pytest.approx(0.2 / np.sqrt(12), rel=0.15)                  # uniform(-0.1, 0.1)


def test_group_environment_reset_structure():
    """sim.group_human (crowd_sim.py:559-622): circles of 4-9 static humans (v_pref 0, goal = position, config radius) whose
    discs do not overlap other groups, at most 4 walkers that keep clear of the groups, robot start and goal on the circle
    of radius 5.5 outside the groups.  (The distributions are pinned against the reference in test_spawn_distribution.py.)"""
    n, H = 512, 13
    cfg_obj = Config(human_num=H)
    cfg_obj.sim.group_human = True
    cfg = abi.flatten_config(cfg_obj, n, phase="train")
    assert cfg.group_human == 1
    st = crowd_oracle.OracleState(n, H)
    crowd_oracle.reset(cfg, st, n_threads=4)
    hum, grp, rob = st.humans.astype(np.float64), st.groups.astype(np.float64), st.robot.astype(np.float64)
    static = hum[:, :, 7] == 0
    n_static = static.sum(1)
    assert ((H - n_static) >= 1).all() and ((H - n_static) <= 4).all()
    assert (static[:, :-1] >= static[:, 1:]).all()                     # the static humans come first
    assert np.all(hum[static][:, 5:7] == hum[static][:, 0:2]) and np.allclose(hum[static][:, 4], 0.3)
    n_groups = (grp[:, :, 3] == 1).sum(1)
    assert (n_groups >= 1).all() and (n_groups <= 3).all()
    for e in range(n):
        k = 0
        for g in range(n_groups[e]):
            r, cx, cy = grp[e, g, 0:3]
            m = int(round(r * 2 * np.pi / 0.6))                        # group radius = 0.3 * 2 * m / (2 pi)
            assert 4 <= m <= 9
            d = np.linalg.norm(hum[e, k:k + m, 0:2] - [cx, cy], axis=1)
            assert np.allclose(d, r, atol=1e-5)
            k += m
            for g2 in range(g):                                        # centres at least r1 + r2 + 2 * 0.3 apart
                assert np.linalg.norm(grp[e, g, 1:3] - grp[e, g2, 1:3]) >= r + grp[e, g2, 0] + 0.6 - 1e-5
        assert k == n_static[e]
        for w in range(k, H):                                          # walkers: outside every group by radius + 1
            for g in range(n_groups[e]):
                assert np.linalg.norm(hum[e, w, 0:2] - grp[e, g, 1:3]) > grp[e, g, 0] + hum[e, w, 4] + 1.0 - 1e-5
    assert np.allclose(np.linalg.norm(rob[:, 0:2], axis=1), 5.5, atol=1e-5)
    assert np.allclose(np.linalg.norm(rob[:, 5:7], axis=1), 5.5, atol=1e-5)
    assert np.all(np.linalg.norm(rob[:, 5:7] - rob[:, 0:2], axis=1) > 5.0)


def test_group_environment_static_humans_stay_put():
    n, H = 128, 9
    cfg_obj = Config(human_num=H)
    cfg_obj.sim.group_human = True
    cfg = abi.flatten_config(cfg_obj, n, phase="train")
    st = crowd_oracle.OracleState(n, H)
    crowd_oracle.reset(cfg, st, n_threads=4)
    p0, static = st.humans[:, :, 0:2].copy(), st.humans[:, :, 7] == 0
    rng = np.random.default_rng(1)
    alive = np.ones(n, bool)
    for _ in range(30):
        out = crowd_oracle.step(cfg, st, rng.normal(0, 0.7, (n, 2)).astype(np.float32), auto_reset=False, n_threads=4)
        alive &= ~out.done.astype(bool)
    keep = alive[:, None] & static
    assert keep.sum() > 100
    assert np.array_equal(st.humans[:, :, 0:2][keep], p0[keep])                     # ORCA with max speed 0: no motion
    moved = np.linalg.norm(st.humans[:, :, 0:2] - p0, axis=-1)[alive[:, None] & ~static]
    assert (moved > 0.5).mean() > 0.8                                                # the walkers walk

"""GPU parity of the crowd-step kernel (K1), through the C ABI:
  * against the golden vectors of the reference's own CrowdSimDict.step (north_star tolerances), and
  * against the C oracle on the same inputs (flags / indices bit-exact, floats to 2e-6),
at fixture sizes and on larger seeded random batches, with goal re-sampling on."""
import numpy as np
import pytest
import torch

from crowdnav_dsrnn_b200 import Config, abi
from oracle import crowd_oracle, state_sampler
from helpers import STEP_CASES, check_against_reference, load_step_case
import gpu_helpers as G

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["group", "thread", "split"])
def phase_a_form(request, monkeypatch):
    """Every case runs with the three forms of the step: one lane group per human in one launch (the default below 4096
    envs), one thread per human (CN_STEP_SEQ=1), and the two-launch form of large batches (ORCA kernel + env-tail kernel,
    CN_STEP_SPLIT=1) -- the switches are read by the launcher at every call, and all forms must give the same bits."""
    monkeypatch.delenv("CN_STEP_SEQ", raising=False)
    monkeypatch.setenv("CN_STEP_SPLIT", "1" if request.param == "split" else "0")
    if request.param == "thread":
        monkeypatch.setenv("CN_STEP_SEQ", "1")
    return request.param


@pytest.mark.parametrize("name", STEP_CASES)
def test_step_matches_reference_golden_and_oracle(name):
    d, cfg_obj, cfg, n = load_step_case(name)
    eng = G.make_engine(cfg_obj, n)
    inp = {f: d["in_" + f] for f in ("robot", "humans", "belief", "extras", "counters")}
    eng.set_state(**inp)
    buf = eng.step(torch.from_numpy(d["in_action"]).cuda(), auto_reset=False)
    out, st = G.buf_to_numpy(buf), G.state_to_numpy(eng.get_state())
    check_against_reference(d, cfg, out, st, where=name)
    ost = G.oracle_state_from(inp, n, cfg.human_num)
    oout = crowd_oracle.step(cfg, ost, d["in_action"], auto_reset=False)
    G.compare(out, st, G.oracle_out_to_numpy(oout), ost, where=name)


def _cfg(human_num, kinematics="holonomic", **over):
    c = Config(kinematics=kinematics, human_num=human_num)
    for k, v in over.items():
        sec, _, attr = k.partition(".")
        setattr(getattr(c, sec), attr, v)
    return c


RANDOM_CASES = [
    ("h5_holo", dict(human_num=5), 4096),
    ("h10_uni_dt01", dict(human_num=10, kinematics="unicycle", **{"env.time_step": 0.1, "reward.discomfort_penalty_factor": 1.0}), 2048),
    ("h20_fov", dict(human_num=20, **{"robot.FOV": 0.5}), 2048),
    ("h1", dict(human_num=1), 512),
    ("h3_robot_visible", dict(human_num=3, **{"robot.visible": True}), 512),
    ("h9_humanfov_uni", dict(human_num=9, kinematics="unicycle", **{"humans.FOV": 1.2}), 512),
    ("h17", dict(human_num=17), 512),
    ("h32", dict(human_num=32), 256),
    ("h31_robot_visible", dict(human_num=31, **{"robot.visible": True}), 256),
    # the optional human behaviours (their own instantiation of the kernel)
    ("n4_social_force_h7", dict(human_num=7, **{"humans.policy": "social_force"}), 1024),
    ("n4_mixed_policies_h12", dict(human_num=12, **{"humans.random_policy_changing": True, "robot.visible": True}), 512),
    ("n4_unobservability_h6", dict(human_num=6, **{"humans.random_unobservability": True, "humans.unobservable_chance": 0.5}), 1024),
    ("n4_random_radii_vpref_h5", dict(human_num=5, **{"humans.random_radii": True, "humans.random_v_pref": True}), 2048),
    ("n4_everything_h20_uni", dict(human_num=20, kinematics="unicycle",
                                   **{"humans.random_policy_changing": True, "humans.random_unobservability": True,
                                      "humans.random_radii": True, "humans.random_v_pref": True, "humans.FOV": 1.5}), 512),
]


@pytest.mark.parametrize("name,kw,n", RANDOM_CASES, ids=[c[0] for c in RANDOM_CASES])
@pytest.mark.parametrize("auto_reset", [False, True])
def test_step_matches_oracle_random(name, kw, n, auto_reset):
    cfg_obj = _cfg(**kw)
    eng = G.make_engine(cfg_obj, n)
    cfg = eng.cfg
    inp = state_sampler.sample(cfg, n, seed=1234 + n)
    act = inp.pop("action")
    eng.set_state(**inp)
    buf = eng.step(torch.from_numpy(act).cuda(), auto_reset=auto_reset)
    out, st = G.buf_to_numpy(buf), G.state_to_numpy(eng.get_state())
    ost = G.oracle_state_from(inp, n, cfg.human_num)
    oout = crowd_oracle.step(cfg, ost, act, auto_reset=auto_reset, n_threads=8)
    rep = G.compare(out, st, G.oracle_out_to_numpy(oout), ost, where=name)
    assert out["done"].sum() > 0 and (out["event"] == abi.EV_DANGER).sum() > 0
    assert rep["bit_exact_humans"] > 0.999, rep


@pytest.mark.parametrize("kw", [dict(human_num=5), dict(human_num=10, kinematics="unicycle"),
                                dict(human_num=20, **{"robot.FOV": 0.5}),
                                dict(human_num=5, **{"test.social_metrics": True, "sim.circle_radius": 4}),
                                dict(human_num=8, **{"sim.group_human": True}),
                                dict(human_num=23, kinematics="unicycle", **{"sim.group_human": True})],
                         ids=["h5", "h10_uni", "h20_fov", "social", "n4_group_h8", "n4_group_h23_uni"])
def test_reset_matches_oracle(kw):
    n = 2048
    cfg_obj = _cfg(**kw)
    eng = G.make_engine(cfg_obj, n, seed=7)
    cfg = eng.cfg
    buf = eng.reset()
    out, st = G.buf_to_numpy(buf), G.state_to_numpy(eng.get_state())
    ost = crowd_oracle.OracleState(n, cfg.human_num)
    oout = crowd_oracle.reset(cfg, ost, n_threads=8)
    for f in ("robot_node", "temporal_edges", "spatial_edges"):
        assert np.abs(out[f] - getattr(oout, f)).max() <= 2e-6, f
    assert np.array_equal(out["visible_mask"], oout.visible_mask.astype(np.int64))
    for f in G.STATE_FIELDS:
        a, b = st[f], getattr(ost, f)
        if f == "counters":
            assert np.array_equal(a, b)
        else:
            assert np.abs(a - b).max() <= 2e-6, f
    # second reset of a subset only touches the masked envs
    mask = (np.arange(n) % 3 == 0).astype(np.uint8)
    eng.reset(torch.from_numpy(mask).cuda())
    crowd_oracle.reset(cfg, ost, mask=mask, n_threads=8)
    st2 = G.state_to_numpy(eng.get_state())
    assert np.array_equal(st2["counters"], ost.counters)
    assert np.abs(st2["humans"] - ost.humans).max() <= 2e-6


@pytest.mark.parametrize("kw,steps", [(dict(human_num=5), 120), (dict(human_num=10, kinematics="unicycle"), 60),
                                      (dict(human_num=20, **{"robot.FOV": 0.5}), 40),
                                      (dict(human_num=8, **{"humans.random_policy_changing": True, "humans.random_unobservability": True,
                                                            "humans.random_radii": True, "humans.random_v_pref": True}), 100),
                                      (dict(human_num=9, **{"sim.group_human": True}), 150)],
                         ids=["h5", "h10_uni", "h20_fov", "n4_options_h8", "n4_group_h9"])
def test_rollout_tracks_oracle(kw, steps):
    """Multi-step trajectory (auto-reset + goal re-sampling on) with a fixed action tape: the CUDA path and the
    oracle must stay in lock-step; flags are compared every step."""
    n = 256
    cfg_obj = _cfg(**kw)
    eng = G.make_engine(cfg_obj, n, seed=3)
    cfg = eng.cfg
    eng.reset()
    ost = crowd_oracle.OracleState(n, cfg.human_num)
    crowd_oracle.reset(cfg, ost, n_threads=8)
    rng = np.random.default_rng(5)
    scale = 0.08 if cfg.kinematics == abi.UNICYCLE else 0.7
    n_done = 0
    for t in range(steps):
        act = rng.normal(0, scale, (n, 2)).astype(np.float32)
        buf = eng.step(torch.from_numpy(act).cuda(), auto_reset=True)
        oout = crowd_oracle.step(cfg, ost, act, auto_reset=True, n_threads=8)
        out = G.buf_to_numpy(buf)
        assert np.array_equal(out["done"], oout.done), "step %d" % t
        assert np.array_equal(out["event"], oout.event), "step %d" % t
        assert np.array_equal(out["visible_mask"], oout.visible_mask.astype(np.int64)), "step %d" % t
        assert np.array_equal(out["goal_changed"], oout.goal_changed.astype(np.int64)), "step %d" % t
        assert np.abs(out["reward"] - oout.reward).max() <= 1e-5
        n_done += int(out["done"].sum())
    st = G.state_to_numpy(eng.get_state())
    assert np.abs(st["humans"] - ost.humans).max() <= 1e-4
    assert np.abs(st["robot"] - ost.robot).max() <= 1e-4
    assert np.abs(st["groups"] - ost.groups).max() <= 2e-6      # circle groups of the running episodes (group environment)
    assert n_done > 0


@pytest.mark.parametrize("refill_threads", ["64", "32", "256"])
def test_spare_and_fallback_resets_track_oracle(refill_threads, monkeypatch):
    """Episodes that end are replaced by their pre-generated spare episode inside the step kernel; envs whose counters
    were changed behind the spares' back (cn_env_set_state) must fall back to the synchronous reset.  Both paths in one
    batch, in lock-step with the oracle's plain reset -- whatever the CTA size of the refill (thread t = try t, t + threads,
    ...: the first accepted try is the same)."""
    monkeypatch.setenv("CN_REFILL_THREADS", refill_threads)
    n = 512
    cfg_obj = _cfg(human_num=5)
    eng = G.make_engine(cfg_obj, n, seed=11)
    cfg = eng.cfg
    eng.reset()
    ost = crowd_oracle.OracleState(n, cfg.human_num)
    crowd_oracle.reset(cfg, ost, n_threads=8)
    st = eng.get_state()
    st["counters"][::2, 2] += 7                      # case_counter of the even envs: their spares are now stale
    st["counters"][::4, 1] += 1                      # scenario_counter of every fourth env as well
    eng.set_state(counters=st["counters"])
    ost.counters[:] = st["counters"].cpu().numpy()
    rng = np.random.default_rng(9)
    n_done = np.zeros(n, dtype=np.int64)
    for t in range(150):
        act = rng.normal(0, 0.7, (n, 2)).astype(np.float32)
        buf = eng.step(torch.from_numpy(act).cuda(), auto_reset=True)
        oout = crowd_oracle.step(cfg, ost, act, auto_reset=True, n_threads=8)
        out = G.buf_to_numpy(buf)
        assert np.array_equal(out["done"], oout.done), "step %d" % t
        assert np.array_equal(out["event"], oout.event), "step %d" % t
        assert np.abs(out["spatial_edges"] - oout.spatial_edges).max() <= 1e-4, "step %d" % t
        n_done += out["done"].astype(np.int64)
    stf = G.state_to_numpy(eng.get_state())
    assert np.array_equal(stf["counters"], ost.counters)
    assert np.abs(stf["humans"] - ost.humans).max() <= 1e-4
    # even envs: fall-back at their first reset, spares afterwards; odd envs: spares only
    assert (n_done[::2] >= 1).sum() > 100 and (n_done[1::2] >= 1).sum() > 100 and (n_done >= 2).sum() > 20

"""SURVEY.md test plan T6 / 8(f) N3 -- the reference's OWN, UNMODIFIED train.py and test.py (run with runpy from
/root/reference) driving this package's `make_vec_envs` boundary object: only the import
`pytorchBaselines.a2c_ppo_acktr.envs.make_vec_envs` is redirected; Policy, PPO, storage,
evaluation.py, the config, the logging / checkpoint / progress.csv code are the reference's.

Build container only (the reference tree is not on the GPU box; there is no GPU here), so the batched simulation
behind the boundary is the C oracle through tests/cpu_engine.OracleEngine -- bit-identical to the CUDA kernels
(tests/test_gpu_crowd_step.py) -- and what is under test is the HOST side of the drop-in: `CrowdVecEnv`, `LazyInfos`
(the dict / event-object contract train.py:263-282 and evaluation.py:155-190 walk), `.venv.envs[0].env`, spaces, auto-reset.
The CUDA twin of the rollout loop is tests/test_gpu_api.py."""
import csv
import os
import runpy
import sys

import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="needs /root/reference (build container only)")


@pytest.fixture()
def reference_with_our_vec_env(monkeypatch):
    ref_import.install_shims()
    import types

    if "torch.utils.tensorboard" not in sys.modules:        # tensorboard is not installed here: inert SummaryWriter
        tb = types.ModuleType("torch.utils.tensorboard")

        class SummaryWriter:
            def __init__(self, *a, **k):
                self.scalars = []

            def add_scalar(self, *a, **k):
                self.scalars.append(a)

        tb.SummaryWriter = SummaryWriter
        monkeypatch.setitem(sys.modules, "torch.utils.tensorboard", tb)
    import crowd_sim  # noqa: F401
    import pytorchBaselines.a2c_ppo_acktr.envs as ref_envs
    from crowd_nav.configs.config import Config

    import crowdnav_dsrnn_b200.envs as our_envs
    from cpu_engine import OracleEngine

    monkeypatch.setattr(our_envs, "CrowdEngine", OracleEngine)             # no GPU here: the oracle steps the batch
    monkeypatch.setattr(ref_envs, "make_vec_envs", our_envs.make_vec_envs)  # THE redirected import
    return Config


def test_unmodified_reference_train_py_runs_on_our_vec_env(reference_with_our_vec_env, monkeypatch, tmp_path):
    Config = reference_with_our_vec_env
    n, updates = 8, 2
    out_dir = str(tmp_path / "run")
    for sec, attr, val in (("training", "num_processes", n), ("training", "num_env_steps", updates * 30 * n), ("training", "cuda", False),
                           ("training", "output_dir", out_dir), ("training", "save_interval", 1), ("training", "log_interval", 1),
                           ("training", "overwrite", False), ("training", "resume", False), ("ppo", "num_mini_batch", 2)):
        monkeypatch.setattr(getattr(Config, sec), attr, val)
    monkeypatch.chdir(ref_import.REFERENCE_ROOT)          # train.py copies ./crowd_nav/configs into the output directory
    monkeypatch.setattr(sys, "argv", ["train.py"])
    torch.set_num_threads(4)
    runpy.run_path(os.path.join(ref_import.REFERENCE_ROOT, "train.py"), run_name="__main__")
    # artefacts of train.py:47-75, 326-404
    assert sorted(os.listdir(os.path.join(out_dir, "checkpoints"))) == ["00000.pt", "00001.pt"]
    assert os.path.exists(os.path.join(out_dir, "configs", "train_config.py")) and os.path.islink(os.path.join(out_dir, "configs", "config.py"))
    assert os.path.exists(os.path.join(out_dir, "output.log"))
    with open(os.path.join(out_dir, "progress.csv")) as f:
        rows = list(csv.DictReader(f))
    assert list(rows[0]) == ["misc/nupdates", "misc/total_timesteps", "fps", "eprewmean", "loss/policy_entropy", "loss/policy_loss",
                             "loss/value_loss"]
    assert int(rows[-1]["misc/total_timesteps"]) == updates * 30 * n
    sd = torch.load(os.path.join(out_dir, "checkpoints", "00001.pt"))
    assert len(sd) == 45 and all(torch.isfinite(v).all() for v in sd.values())


def test_unmodified_reference_test_py_runs_on_our_vec_env(reference_with_our_vec_env, monkeypatch, tmp_path):
    Config = reference_with_our_vec_env
    import shutil

    model_dir = tmp_path / "data" / "m"
    (model_dir / "checkpoints").mkdir(parents=True)
    shutil.copy(os.path.join(ref_import.REFERENCE_ROOT, "data/example_model/checkpoints/27776.pt"), model_dir / "checkpoints" / "27776.pt")
    episodes = 40
    monkeypatch.setattr(Config.env, "test_size", episodes)
    monkeypatch.setattr(Config.training, "cuda", False)
    monkeypatch.chdir(tmp_path)                            # test.py resolves --model_dir against the working directory
    monkeypatch.setattr(sys, "argv", ["test.py", "--model_dir", "data/m", "--test_model", "27776.pt", "--test_name", "dropin"])
    torch.set_num_threads(2)
    import logging

    root = logging.getLogger()
    for h in list(root.handlers):                          # test.py configures logging with basicConfig: start from a clean root logger
        root.removeHandler(h)
    runpy.run_path(os.path.join(ref_import.REFERENCE_ROOT, "test.py"), run_name="__main__")
    for h in list(root.handlers):
        h.flush()
    log = (model_dir / "test" / "model_27776_test_dropin_.log").read_text()      # name built at test.py:103-128
    assert "Using model" in log and "success rate" in log.lower()
    # the shipped policy solves most of the 40 test cases on the batched backend too (reference: 0.93 over 2000 episodes)
    import re

    rates = [re.search(r"%s rate: ([0-9.]+)" % k, log) for k in ("success", "collision", "timeout")]      # evaluation.py:283-300
    assert all(r is not None for r in rates), log[-2000:]
    success, collision, timeout = (float(r.group(1)) for r in rates)
    assert abs(success + collision + timeout - 1.0) < 2e-3 and success >= 0.7

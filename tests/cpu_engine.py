"""TEST INFRASTRUCTURE -- a stand-in for `crowdnav_dsrnn_b200.engine.CrowdEngine` backed by the C oracle, with the same
methods and `StepBuffers` outputs (CPU tensors).  It exists so that the HOST side of the drop-in boundary (envs.py:
`make_vec_envs`, `CrowdVecEnv`, `LazyInfos`, `_EnvView`) can be driven by the reference's own unmodified train.py / test.py
in the build container, where there is no GPU (tests/test_reference_drivers.py).  The product never imports this."""
import numpy as np
import torch

from crowdnav_dsrnn_b200 import abi
from crowdnav_dsrnn_b200.engine import StepBuffers
from oracle import crowd_oracle


class OracleEngine:
    def __init__(self, config, n_envs, device, phase=None, seed=None, env_id_offset=0, nenv=None, **tries):
        self.device = torch.device("cpu")
        self.n, self.h = int(n_envs), int(config.sim.human_num)
        self.config = config
        self.cfg = abi.flatten_config(config, n_envs, phase=phase, seed=seed, env_id_offset=env_id_offset, nenv=nenv, **tries)
        self.state = crowd_oracle.OracleState(self.n, self.h)
        self.bufs = [StepBuffers(self.n, self.h, self.device), StepBuffers(self.n, self.h, self.device)]
        self.cur = 0
        self.launches = 0
        self.handle = None

    def _fill(self, b, out, obs_only=False):
        names = ("robot_node", "temporal_edges", "spatial_edges") if obs_only else \
            ("robot_node", "temporal_edges", "spatial_edges", "reward", "done", "event", "scenario", "info", "episode_return", "episode_length")
        for k in names:
            getattr(b, k).copy_(torch.from_numpy(np.asarray(getattr(out, k))).view_as(getattr(b, k)))
        if not obs_only:
            b.not_done.copy_((1.0 - torch.from_numpy(out.done.astype(np.float32))).view(self.n, 1))
        return b

    def reset(self, mask=None):
        self.cur ^= 1
        out = crowd_oracle.reset(self.cfg, self.state, None if mask is None else np.asarray(mask, np.uint8))
        return self._fill(self.bufs[self.cur], out, obs_only=True)

    def step(self, action, auto_reset=True, defer_refill=False):
        self.cur ^= 1
        out = crowd_oracle.step(self.cfg, self.state, action.detach().cpu().numpy(), auto_reset=auto_reset)
        return self._fill(self.bufs[self.cur], out)

    def join(self):
        pass

    def get_state(self):
        return {k: torch.from_numpy(v.copy()) for k, v in self.state.as_dict().items()}

    def set_state(self, **fields):
        for k, v in fields.items():
            if v is not None:
                getattr(self.state, k)[...] = np.asarray(v)

    def close(self):
        pass

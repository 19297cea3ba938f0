"""crowdnav_dsrnn_b200 -- B200-native (sm_100a) rollout hot path of CrowdNav-DSRNN.

Crowd step + device reset + DS-RNN policy forward as hand-written CUDA behind
the reference's own Python API (CrowdSimDict / make_vec_envs / Policy).
The CUDA library (csrc/ -> libcrowdnav_b200.so) is loaded on first use and
there is no CPU fallback: every compute entry point raises if it is missing.
"""
from .config import Config, BaseConfig  # noqa: F401

__all__ = ["Config", "BaseConfig"]

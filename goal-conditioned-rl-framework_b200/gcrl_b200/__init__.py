"""gcrl_b200 -- B200-native (sm_100a) HER-sample + off-policy-update hot path.

Host-side mirror of the reference's duck-typed agent/buffer interface
(CodeKnight314/Goal-Conditioned-RL-Framework, src/agent.py, src/buffer.py,
src/utils.py) over the C-ABI library ``libgcrl_b200.so`` (include/gcrl_b200.h).
There is no CPU fallback: importing this package without the built library
raises, and every compute call needs a CUDA device.
"""
from ._lib import GcrlError, lib, library_path  # noqa: F401
from .buffer import HERBuffer  # noqa: F401
from .replay import PERBuffer, ReplayBuffer  # noqa: F401
from .normalizer import RunningNormalizer  # noqa: F401
from .agent import DDPG, TD3Agent, CosineAnnealingLR  # noqa: F401
from .sac import SACAgent, TQCAgent  # noqa: F401

__all__ = ["HERBuffer", "PERBuffer", "ReplayBuffer", "RunningNormalizer", "DDPG", "TD3Agent", "SACAgent", "TQCAgent", "GcrlError", "lib",
           "library_path", "CosineAnnealingLR"]

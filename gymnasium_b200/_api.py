"""The host-framework types the engine plugs into: Gymnasium's when it is importable, a minimal stand-in otherwise.

With Gymnasium present, ``B200VectorEnv`` IS a ``gymnasium.vector.VectorEnv`` (required by vector wrappers,
gymnasium/vector/vector_env.py:374-377) and is created by ``gymnasium.make_vec`` through a spec's
``vector_entry_point`` (gymnasium/envs/registration.py:933-963).
"""
from __future__ import annotations

import os as _os

try:
    if _os.environ.get("B200ENV_FORCE_COMPAT"):
        raise ImportError("B200ENV_FORCE_COMPAT is set")
    import gymnasium as _gym
    from gymnasium.spaces import Box, Discrete, MultiDiscrete, Tuple
    from gymnasium.vector import AutoresetMode, VectorEnv
    from gymnasium.vector.utils import batch_space

    HAVE_GYMNASIUM = True
    gymnasium = _gym
except ImportError:  # pragma: no cover - exercised on boxes without the host framework
    from ._compat import AutoresetMode, Box, Discrete, MultiDiscrete, Tuple, VectorEnv, batch_space

    HAVE_GYMNASIUM = False
    gymnasium = None

__all__ = ["AutoresetMode", "Box", "Discrete", "MultiDiscrete", "Tuple", "VectorEnv", "batch_space", "HAVE_GYMNASIUM",
           "gymnasium"]

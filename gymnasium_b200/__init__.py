"""gymnasium_b200 -- a B200-native (sm_100a) vectorised step()/reset() engine behind Gymnasium's VectorEnv API.

One fused CUDA launch per ``step()`` for a struct-of-arrays batch of N environments; see DESIGN.md.
Importing the package does not need a GPU; constructing an environment does (there is no CPU fallback).
"""
from ._api import HAVE_GYMNASIUM, AutoresetMode
from .registration import install, make_vec, register_envs, uninstall
from .vector_env import B200VectorEnv

__version__ = "0.1.0"
__all__ = ["AutoresetMode", "B200VectorEnv", "HAVE_GYMNASIUM", "install", "make_vec", "register_envs", "uninstall",
           "CartPoleVectorEnv", "CliffWalkingVectorEnv", "FrozenLakeVectorEnv", "HumanoidVectorEnv", "LunarLanderVectorEnv",
           "TaxiVectorEnv", "AcrobotVectorEnv", "MountainCarVectorEnv", "MountainCarContinuousVectorEnv",
           "PendulumVectorEnv"]

register_envs()


def __getattr__(name):
    if name in ("CartPoleVectorEnv", "CliffWalkingVectorEnv", "FrozenLakeVectorEnv", "HumanoidVectorEnv",
                "LunarLanderVectorEnv", "TaxiVectorEnv", "AcrobotVectorEnv", "MountainCarVectorEnv",
                "MountainCarContinuousVectorEnv", "PendulumVectorEnv"):
        from . import envs

        return getattr(envs, name)
    raise AttributeError(name)

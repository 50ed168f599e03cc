from .cartpole import CartPoleVectorEnv
from .frozen_lake import FrozenLakeVectorEnv
from .lunar_lander import LunarLanderVectorEnv

__all__ = ["CartPoleVectorEnv", "FrozenLakeVectorEnv", "LunarLanderVectorEnv"]

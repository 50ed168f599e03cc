from .cartpole import CartPoleVectorEnv
from .frozen_lake import FrozenLakeVectorEnv

__all__ = ["CartPoleVectorEnv", "FrozenLakeVectorEnv"]

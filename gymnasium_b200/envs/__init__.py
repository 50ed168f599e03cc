from .cartpole import CartPoleVectorEnv
from .frozen_lake import FrozenLakeVectorEnv
from .humanoid import HumanoidVectorEnv
from .lunar_lander import LunarLanderVectorEnv

__all__ = ["CartPoleVectorEnv", "FrozenLakeVectorEnv", "HumanoidVectorEnv", "LunarLanderVectorEnv"]

from .cartpole import CartPoleVectorEnv
from .frozen_lake import FrozenLakeVectorEnv, TabularVectorEnv
from .humanoid import HumanoidVectorEnv
from .lunar_lander import LunarLanderVectorEnv
from .toy_text import CliffWalkingVectorEnv, TaxiVectorEnv

__all__ = ["CartPoleVectorEnv", "CliffWalkingVectorEnv", "FrozenLakeVectorEnv", "HumanoidVectorEnv",
           "LunarLanderVectorEnv", "TabularVectorEnv", "TaxiVectorEnv"]

from .cartpole import CartPoleVectorEnv
from .classic_control import (AcrobotVectorEnv, MountainCarContinuousVectorEnv, MountainCarVectorEnv,
                              PendulumVectorEnv)
from .frozen_lake import FrozenLakeVectorEnv, TabularVectorEnv
from .hopper import HopperVectorEnv, Walker2dVectorEnv
from .humanoid import HumanoidVectorEnv
from .lunar_lander import LunarLanderVectorEnv
from .toy_text import CliffWalkingVectorEnv, TaxiVectorEnv

__all__ = ["AcrobotVectorEnv", "CartPoleVectorEnv", "CliffWalkingVectorEnv", "FrozenLakeVectorEnv", "HopperVectorEnv", "HumanoidVectorEnv",
           "LunarLanderVectorEnv", "MountainCarContinuousVectorEnv", "MountainCarVectorEnv", "PendulumVectorEnv",
           "TabularVectorEnv", "TaxiVectorEnv", "Walker2dVectorEnv"]

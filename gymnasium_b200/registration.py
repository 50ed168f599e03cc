"""Plugging the engine into Gymnasium's registry (gymnasium/envs/registration.py).

Three ways in, all ending in the same ``B200VectorEnv`` subclasses:

* ``gymnasium.make_vec("B200/CartPole-v1", num_envs=N)`` -- ids registered by :func:`register_envs` (called on import)
  whose ``vector_entry_point`` is the engine (registration.py:933-963) and whose ``entry_point`` is still the reference's
  single env, so ``gymnasium.make`` keeps working.
* :func:`install` -- points the ``vector_entry_point`` of the stock ids (``CartPole-v1``, ``FrozenLake-v1``, ...) at the
  engine, so unmodified user code calling ``gymnasium.make_vec("CartPole-v1", num_envs=N)`` lands on the GPU
  (``make_vec`` prefers a spec's vector entry point when no mode is given, registration.py:887-891).
* :func:`make_vec` below -- works without Gymnasium installed.
"""
from __future__ import annotations

from ._api import HAVE_GYMNASIUM, gymnasium

# id -> (vector entry point, reference entry point, max_episode_steps, reward_threshold, default kwargs)
ENVS = {
    "CartPole-v1": ("gymnasium_b200.envs.cartpole:CartPoleVectorEnv",
                    "gymnasium.envs.classic_control.cartpole:CartPoleEnv", 500, 475.0, {}),
    "CartPole-v0": ("gymnasium_b200.envs.cartpole:CartPoleVectorEnv",
                    "gymnasium.envs.classic_control.cartpole:CartPoleEnv", 200, 195.0, {}),
    "FrozenLake-v1": ("gymnasium_b200.envs.frozen_lake:FrozenLakeVectorEnv",
                      "gymnasium.envs.toy_text.frozen_lake:FrozenLakeEnv", 100, 0.70, {"map_name": "4x4"}),
    "FrozenLake8x8-v1": ("gymnasium_b200.envs.frozen_lake:FrozenLakeVectorEnv",
                         "gymnasium.envs.toy_text.frozen_lake:FrozenLakeEnv", 200, 0.85, {"map_name": "8x8"}),
    "MountainCar-v0": ("gymnasium_b200.envs.classic_control:MountainCarVectorEnv",
                       "gymnasium.envs.classic_control.mountain_car:MountainCarEnv", 200, -110.0, {}),
    "MountainCarContinuous-v0": ("gymnasium_b200.envs.classic_control:MountainCarContinuousVectorEnv",
                                 "gymnasium.envs.classic_control.continuous_mountain_car:Continuous_MountainCarEnv",
                                 999, 90.0, {}),
    "Pendulum-v1": ("gymnasium_b200.envs.classic_control:PendulumVectorEnv",
                    "gymnasium.envs.classic_control.pendulum:PendulumEnv", 200, None, {}),
    "Acrobot-v1": ("gymnasium_b200.envs.classic_control:AcrobotVectorEnv",
                   "gymnasium.envs.classic_control.acrobot:AcrobotEnv", 500, -100.0, {}),
    "CliffWalking-v1": ("gymnasium_b200.envs.toy_text:CliffWalkingVectorEnv",
                        "gymnasium.envs.toy_text.cliffwalking:CliffWalkingEnv", None, None, {}),
    "CliffWalkingSlippery-v1": ("gymnasium_b200.envs.toy_text:CliffWalkingVectorEnv",
                                "gymnasium.envs.toy_text.cliffwalking:CliffWalkingEnv", None, None, {"is_slippery": True}),
    "Blackjack-v1": ("gymnasium_b200.envs.blackjack:BlackjackVectorEnv", "gymnasium.envs.toy_text.blackjack:BlackjackEnv",
                     None, None, {"sab": True, "natural": False}),
    "Taxi-v4": ("gymnasium_b200.envs.toy_text:TaxiVectorEnv", "gymnasium.envs.toy_text.taxi:TaxiEnv", 200, 8, {}),
    "LunarLander-v3": ("gymnasium_b200.envs.lunar_lander:LunarLanderVectorEnv",
                       "gymnasium.envs.box2d.lunar_lander:LunarLander", 1000, 200, {}),
    "LunarLanderContinuous-v3": ("gymnasium_b200.envs.lunar_lander:LunarLanderVectorEnv",
                                 "gymnasium.envs.box2d.lunar_lander:LunarLander", 1000, 200, {"continuous": True}),
    "Humanoid-v5": ("gymnasium_b200.envs.humanoid:HumanoidVectorEnv",
                    "gymnasium.envs.mujoco.humanoid_v5:HumanoidEnv", 1000, None, {}),
    "Hopper-v5": ("gymnasium_b200.envs.hopper:HopperVectorEnv", "gymnasium.envs.mujoco.hopper_v5:HopperEnv", 1000, 3800.0, {}),
    "Walker2d-v5": ("gymnasium_b200.envs.hopper:Walker2dVectorEnv", "gymnasium.envs.mujoco.walker2d_v5:Walker2dEnv", 1000,
                    None, {}),
    "HalfCheetah-v5": ("gymnasium_b200.envs.hopper:HalfCheetahVectorEnv",
                       "gymnasium.envs.mujoco.half_cheetah_v5:HalfCheetahEnv", 1000, 4800.0, {}),
    "InvertedPendulum-v5": ("gymnasium_b200.envs.hopper:InvertedPendulumVectorEnv",
                            "gymnasium.envs.mujoco.inverted_pendulum_v5:InvertedPendulumEnv", 1000, 950.0, {}),
}
NAMESPACE = "B200"


def _load(entry: str):
    import importlib

    mod, attr = entry.split(":")
    return getattr(importlib.import_module(mod), attr)


def register_envs() -> list[str]:
    """Register ``B200/<id>`` specs with Gymnasium (no-op without Gymnasium)."""
    if not HAVE_GYMNASIUM:
        return []
    ids = []
    for env_id, (vec, ref, limit, thr, kwargs) in ENVS.items():
        full = f"{NAMESPACE}/{env_id}"
        if full not in gymnasium.registry:
            gymnasium.register(id=full, entry_point=ref, vector_entry_point=vec, max_episode_steps=limit,
                               reward_threshold=thr, kwargs=dict(kwargs))
        ids.append(full)
    return ids


def install(ids=None) -> list[str]:
    """Route the stock ids' ``make_vec`` to the engine by replacing their ``vector_entry_point``."""
    if not HAVE_GYMNASIUM:
        raise RuntimeError("gymnasium is not importable; use gymnasium_b200.make_vec instead")
    done = []
    for env_id in (ids or ENVS):
        spec = gymnasium.registry.get(env_id)
        if spec is None:
            continue
        spec.vector_entry_point = ENVS[env_id][0]
        done.append(env_id)
    return done


def uninstall() -> None:
    """Restore the reference's own vector entry points (CartPole's NumPy CartPoleVectorEnv, none for FrozenLake)."""
    if not HAVE_GYMNASIUM:
        return
    stock = {"CartPole-v1": "gymnasium.envs.classic_control.cartpole:CartPoleVectorEnv",
             "CartPole-v0": "gymnasium.envs.classic_control.cartpole:CartPoleVectorEnv"}
    for env_id in ENVS:
        spec = gymnasium.registry.get(env_id)
        if spec is not None:
            spec.vector_entry_point = stock.get(env_id)


def make_vec(env_id: str, num_envs: int = 1, **kwargs):
    """Standalone equivalent of ``gymnasium.make_vec(id, num_envs, vectorization_mode="vector_entry_point", **kwargs)``."""
    key = env_id.split("/", 1)[1] if env_id.startswith(NAMESPACE + "/") else env_id
    if key not in ENVS:
        raise KeyError(f"{env_id!r} is not a gymnasium_b200 environment; known: {sorted(ENVS)}")
    vec, _, limit, _, defaults = ENVS[key]
    kw = {**defaults, **kwargs}
    kw.setdefault("max_episode_steps", limit)
    return _load(vec)(num_envs=num_envs, **kw)

"""Oracle: CartPole-v1 dynamics (numpy, float64 state), batched over lanes.

Oracle only (see oracle/__init__.py).  Restates
  * ``CartPoleEnv.__init__`` constants  gymnasium/envs/classic_control/cartpole.py:124-136
  * ``CartPoleEnv.step``   cartpole.py:164-226 -- explicit Euler (``kinematics_integrator == "euler"`` :185-189),
    strict thresholds :198-203, reward 1.0 also on the terminating step :205-211
    (``sutton_barto_reward``: 0 / -1)
  * ``CartPoleEnv.reset``  cartpole.py:228-247 -- ``np_random.uniform(low, high, size=(4,))`` with
    ``maybe_parse_reset_bounds`` gymnasium/envs/classic_control/utils.py:17-46 (defaults +-0.05)
Pinned by tests/golden/cartpole_*.npz (live reference) and the doctest literals at
gymnasium/vector/vector_env.py:154-201 and cartpole.py:84-85.
"""
from __future__ import annotations

import math

import numpy as np

from .vector import OracleVectorEnv

GRAVITY = 9.8
MASSCART = 1.0
MASSPOLE = 0.1
TOTAL_MASS = MASSPOLE + MASSCART
LENGTH = 0.5
POLEMASS_LENGTH = MASSPOLE * LENGTH
FORCE_MAG = 10.0
TAU = 0.02
THETA_THRESHOLD = 12 * 2 * math.pi / 360
X_THRESHOLD = 2.4


def parse_reset_bounds(options, default_low=-0.05, default_high=0.05):
    """classic_control/utils.py:17-46."""
    if options is None:
        return default_low, default_high
    low = options.get("low") if "low" in options else default_low
    high = options.get("high") if "high" in options else default_high
    try:
        low, high = float(low), float(high)
    except (ValueError, TypeError) as e:
        raise ValueError(f"An option ({low}, {high}) could not be converted to a float.") from e
    if low > high:
        raise ValueError(f"Lower bound ({low}) must be lower than higher bound ({high}).")
    return low, high


def cartpole_step_f64(state: np.ndarray, action: np.ndarray) -> np.ndarray:
    """One Euler step on a ``(L, 4)`` float64 state; op order follows cartpole.py:169-189 literally."""
    x, x_dot, theta, theta_dot = state[:, 0], state[:, 1], state[:, 2], state[:, 3]
    force = np.where(action == 1, FORCE_MAG, -FORCE_MAG)
    costheta = np.cos(theta)
    sintheta = np.sin(theta)
    temp = (force + POLEMASS_LENGTH * np.square(theta_dot) * sintheta) / TOTAL_MASS
    thetaacc = (GRAVITY * sintheta - costheta * temp) / (
        LENGTH * (4.0 / 3.0 - MASSPOLE * np.square(costheta) / TOTAL_MASS)
    )
    xacc = temp - POLEMASS_LENGTH * thetaacc * costheta / TOTAL_MASS
    out = np.empty_like(state)
    out[:, 0] = x + TAU * x_dot
    out[:, 1] = x_dot + TAU * xacc
    out[:, 2] = theta + TAU * theta_dot
    out[:, 3] = theta_dot + TAU * thetaacc
    return out


class OracleCartPole(OracleVectorEnv):
    def __init__(self, num_envs, max_episode_steps=500, sutton_barto_reward=False, autoreset_mode="NextStep"):
        super().__init__(num_envs, max_episode_steps, autoreset_mode)
        self.sutton_barto_reward = sutton_barto_reward
        self.state = np.zeros((num_envs, 4), dtype=np.float64)
        self.beyond = np.zeros(num_envs, dtype=np.bool_)  # steps_beyond_terminated is not None (cartpole.py:207-220)

    def _reset_env(self, i, options):
        low, high = parse_reset_bounds(options)
        self.state[i] = self._rng(i).uniform(low=low, high=high, size=(4,))
        self.beyond[i] = False  # cartpole.py:243

    def _step_lanes(self, lanes, actions):
        if np.any((actions < 0) | (actions > 1)):
            raise AssertionError(f"invalid action in {actions!r}")
        s = cartpole_step_f64(self.state[lanes], actions)
        self.state[lanes] = s
        x, theta = s[:, 0], s[:, 2]
        term = (x < -X_THRESHOLD) | (x > X_THRESHOLD) | (theta < -THETA_THRESHOLD) | (theta > THETA_THRESHOLD)
        if self.sutton_barto_reward:
            reward = np.where(term, -1.0, 0.0)
        else:  # 1.0, also on the terminating step; 0.0 on steps taken after it without a reset (DISABLED autoreset only)
            reward = np.where(term & self.beyond[lanes], 0.0, 1.0)
        self.beyond[lanes] |= term
        return reward, term, {}

    def _obs(self):
        return self.state.astype(np.float32)

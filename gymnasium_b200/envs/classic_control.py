"""MountainCar-v0, MountainCarContinuous-v0, Pendulum-v1 and Acrobot-v1 on the B200 engine (``csrc/classic.cu``).

Mirrors ``MountainCarEnv`` (gymnasium/envs/classic_control/mountain_car.py:16-170), ``Continuous_MountainCarEnv``
(continuous_mountain_car.py:26-196), ``PendulumEnv`` (pendulum.py:16-165) and ``AcrobotEnv`` (acrobot.py:29-283) behind the
vector API with SyncVectorEnv's conventions (seed+i PCG64 streams, float64 rewards, TimeLimit, autoreset modes).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _lib
from .._api import AutoresetMode, Box, Discrete
from ..vector_env import B200VectorEnv, ptr
from .cartpole import _parse_reset_bounds

MOUNTAINCAR, MOUNTAINCAR_CONTINUOUS, PENDULUM, ACROBOT = 0, 1, 2, 3


class ClassicVectorEnv(B200VectorEnv):
    metadata = {"render_modes": [], "render_fps": 30, "autoreset_mode": AutoresetMode.NEXT_STEP}
    family: int
    state_size: int
    obs_size: int

    def _init_state(self, params):
        n, dev = self.num_envs, self.device
        self._state = torch.zeros((self.state_size, n), dtype=torch.float64, device=dev)
        self._sflag = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._params = list(params) + [0.0] * (4 - len(params))
        self._cfg = _lib.ClassicCfg(family=self.family, state=self._state.data_ptr(), sflag=self._sflag.data_ptr(),
                                    ctrl=self._ctrl.data_ptr(), rng=ptr(self._rng))
        for k in range(4):
            self._cfg.p[k] = float(self._params[k])

    @property
    def state(self) -> torch.Tensor:
        return self._state.t().contiguous()

    def set_state(self, state) -> None:
        """Overwrites the per-env float64 state from a ``(N, state_size)`` array."""
        t = torch.as_tensor(state, dtype=torch.float64).to(self.device).reshape(self.num_envs, self.state_size)
        self._state.copy_(t.t())

    def _alloc_outputs(self):
        n = self.num_envs
        layout = {"obs": ((n, self.obs_size), torch.float32), "reward": ((n,), torch.float64),
                  "terminated": ((n,), torch.bool), "truncated": ((n,), torch.bool)}
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            layout["final_obs"] = ((n, self.obs_size), torch.float32)
        out = self._alloc_packed(layout)
        if "final_obs" in out:
            out["final_obs"].zero_()
        return out

    def _reset_params(self, options):
        return self._params

    def _reset_kernel(self, mask, options, out):
        p = self._reset_params(options)
        for k in range(4):
            self._cfg.p[k] = float(p[k])
        if mask is not None and self.copy and self._has_reset:
            out["obs"].copy_(self._last_obs)
        _lib.check(
            self._lib.b2e_classic_reset(C.byref(self._batch), C.byref(self._cfg),
                                        ptr(None if mask is None else mask.view(torch.uint8)), ptr(out["obs"]), self._stream),
            "b2e_classic_reset",
        )
        for k in range(4):  # autoreset calls env.reset() without options: back to the defaults
            self._cfg.p[k] = float(self._params[k])
        self._last_obs = out["obs"]

    def _step_kernel(self, actions, out):
        _lib.check(
            self._lib.b2e_classic_step(C.byref(self._batch), C.byref(self._cfg), ptr(actions), ptr(out["obs"]),
                                       ptr(out["reward"]), ptr(out["terminated"]), ptr(out["truncated"]),
                                       ptr(out.get("final_obs")), self._stream),
            "b2e_classic_step",
        )
        self._last_obs = out["obs"]

    def _step_info(self, out):
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            done = out["terminated"] | out["truncated"]
            return {"final_obs": out["final_obs"], "_final_obs": done, "final_info": {}, "_final_info": done}
        return {}

    def _prepare_box_actions(self, actions):
        """(N, 1) float32 torque / force arrays."""
        n = self.num_envs
        if isinstance(actions, torch.Tensor):
            if tuple(actions.shape) != (n, 1):
                raise ValueError(f"expected actions of shape ({n}, 1), got {tuple(actions.shape)}")
            return actions.to(device=self.device, dtype=torch.float32).contiguous()
        a = np.asarray(actions)
        if a.ndim == 0:
            raise TypeError(f"actions must have a leading dimension of num_envs={n}, got a scalar")
        if a.shape != (n, 1):
            raise ValueError(f"expected actions of shape ({n}, 1), got {a.shape}")
        return super()._prepare_actions(np.ascontiguousarray(a, dtype=np.float32))


class MountainCarVectorEnv(ClassicVectorEnv):
    family, state_size, obs_size = MOUNTAINCAR, 2, 2

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 200, goal_velocity: float = 0,
                 render_mode: str | None = None, **engine_kwargs):
        low = np.array([-1.2, -0.07], dtype=np.float32)   # mountain_car.py:118-119
        high = np.array([0.6, 0.07], dtype=np.float32)
        super().__init__(num_envs, Box(low, high, dtype=np.float32), Discrete(3), max_episode_steps=max_episode_steps,
                         render_mode=render_mode, **engine_kwargs)
        self._init_state([-0.6, -0.4, float(goal_velocity)])

    def _reset_params(self, options):
        low, high = _parse_reset_bounds(options, -0.6, -0.4)  # mountain_car.py:162
        return [low, high, self._params[2], 0.0]


class MountainCarContinuousVectorEnv(ClassicVectorEnv):
    family, state_size, obs_size = MOUNTAINCAR_CONTINUOUS, 2, 2
    discrete_actions = False

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 999, goal_velocity: float = 0,
                 render_mode: str | None = None, **engine_kwargs):
        low = np.array([-1.2, -0.07], dtype=np.float32)
        high = np.array([0.6, 0.07], dtype=np.float32)
        act = Box(low=-1.0, high=1.0, shape=(1,), dtype=np.float32)  # continuous_mountain_car.py:139-141
        super().__init__(num_envs, Box(low, high, dtype=np.float32), act, max_episode_steps=max_episode_steps,
                         render_mode=render_mode, **engine_kwargs)
        self._init_state([-0.6, -0.4, float(goal_velocity)])

    def _reset_params(self, options):
        low, high = _parse_reset_bounds(options, -0.6, -0.4)
        return [low, high, self._params[2], 0.0]

    _prepare_actions = ClassicVectorEnv._prepare_box_actions


class PendulumVectorEnv(ClassicVectorEnv):
    family, state_size, obs_size = PENDULUM, 2, 3
    discrete_actions = False

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 200, g: float = 10.0,
                 render_mode: str | None = None, **engine_kwargs):
        high = np.array([1.0, 1.0, 8.0], dtype=np.float32)  # pendulum.py:117-125
        act = Box(low=-2.0, high=2.0, shape=(1,), dtype=np.float32)
        super().__init__(num_envs, Box(low=-high, high=high, dtype=np.float32), act, max_episode_steps=max_episode_steps,
                         render_mode=render_mode, **engine_kwargs)
        self._init_state([math.pi, 1.0, float(g)])

    def _reset_params(self, options):
        if options is None:
            return self._params
        out = []
        for key, default in (("x_init", math.pi), ("y_init", 1.0)):  # pendulum.py:151-158
            x = options.get(key) if key in options else default
            try:
                out.append(float(x))
            except (ValueError, TypeError) as e:
                raise ValueError(f"An option ({x}) could not be converted to a float.") from e
        return [out[0], out[1], self._params[2], 0.0]

    _prepare_actions = ClassicVectorEnv._prepare_box_actions


class AcrobotVectorEnv(ClassicVectorEnv):
    family, state_size, obs_size = ACROBOT, 4, 6
    metadata = {"render_modes": [], "render_fps": 15, "autoreset_mode": AutoresetMode.NEXT_STEP}

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 500, render_mode: str | None = None,
                 **engine_kwargs):
        high = np.array([1.0, 1.0, 1.0, 1.0, 4 * math.pi, 9 * math.pi], dtype=np.float32)  # acrobot.py:170-174
        super().__init__(num_envs, Box(low=-high, high=high, dtype=np.float32), Discrete(3),
                         max_episode_steps=max_episode_steps, render_mode=render_mode, **engine_kwargs)
        self._init_state([-0.1, 0.1])

    def _reset_params(self, options):
        low, high = _parse_reset_bounds(options, -0.1, 0.1)  # acrobot.py:181-185
        return [low, high, 0.0, 0.0]

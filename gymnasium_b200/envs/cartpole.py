"""CartPole-v1 on the B200 engine.

Mirrors ``CartPoleEnv`` (gymnasium/envs/classic_control/cartpole.py:20-352) behind the vector API, the way the in-tree
``CartPoleVectorEnv`` (:355-604) does, but with SyncVectorEnv's conventions (per-env ``seed+i`` PCG64 streams, float64
rewards) and the dynamics in ``gymnasium_b200/csrc/cartpole.cu``.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _lib
from .._api import AutoresetMode, Box, Discrete
from ..vector_env import B200VectorEnv, ptr


def _parse_reset_bounds(options, default_low, default_high):
    """``maybe_parse_reset_bounds`` (gymnasium/envs/classic_control/utils.py:17-46): same checks, same messages."""
    if options is None:
        return default_low, default_high
    low = options.get("low") if "low" in options else default_low
    high = options.get("high") if "high" in options else default_high
    out = []
    for x in (low, high):
        try:
            out.append(float(x))
        except (ValueError, TypeError) as e:
            raise ValueError(f"An option ({x}) could not be converted to a float.") from e
    low, high = out
    if low > high:
        raise ValueError(f"Lower bound ({low}) must be lower than higher bound ({high}).")
    return low, high


class CartPoleVectorEnv(B200VectorEnv):
    """N CartPole-v1 envs, one fused step+autoreset launch per ``step()``.

    Observation ``(N, 4) float32``, action ``(N,) int64`` in {0, 1}, reward ``(N,) float64`` -- the dtypes
    ``SyncVectorEnv`` returns for ``CartPole-v1``.
    """

    metadata = {"render_modes": [], "render_fps": 50, "autoreset_mode": AutoresetMode.NEXT_STEP}

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 500, sutton_barto_reward: bool = False,
                 render_mode: str | None = None, **engine_kwargs):
        x_threshold = 2.4
        theta_threshold = 12 * 2 * math.pi / 360  # cartpole.py:135
        high = np.array([x_threshold * 2, np.inf, theta_threshold * 2, np.inf], dtype=np.float32)  # :140-148
        super().__init__(num_envs, Box(-high, high, dtype=np.float32), Discrete(2),
                         max_episode_steps=max_episode_steps, render_mode=render_mode, **engine_kwargs)
        self.sutton_barto_reward = bool(sutton_barto_reward)
        self._cfg = _lib.CartPoleCfg(reset_low=-0.05, reset_high=0.05, sutton_barto_reward=int(self.sutton_barto_reward))
        # struct of arrays: x[n], x_dot[n], theta[n], theta_dot[n]
        self._state = torch.zeros((4, self.num_envs), dtype=torch.float64, device=self.device)

    # -- buffers -------------------------------------------------------------------------------------------------
    def _alloc_outputs(self):
        n = self.num_envs
        layout = {"obs": ((n, 4), torch.float32), "reward": ((n,), torch.float64),
                  "terminated": ((n,), torch.bool), "truncated": ((n,), torch.bool)}
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            layout["final_obs"] = ((n, 4), torch.float32)
        out = self._alloc_packed(layout)
        if "final_obs" in out:
            out["final_obs"].zero_()
        return out

    @property
    def state(self) -> torch.Tensor:
        """float64 ``(N, 4)`` view-copy of (x, x_dot, theta, theta_dot) -- ``CartPoleEnv.state`` per env."""
        return self._state.t().contiguous()

    def set_state(self, state) -> None:
        """Overwrites the per-env state from a ``(N, 4)`` array of (x, x_dot, theta, theta_dot)."""
        t = torch.as_tensor(state, dtype=torch.float64).to(self.device).reshape(self.num_envs, 4)
        self._state.copy_(t.t())

    # -- kernels -------------------------------------------------------------------------------------------------
    def _reset_kernel(self, mask, options, out):
        low, high = _parse_reset_bounds(options, -0.05, 0.05)  # cartpole.py:236-241
        self._cfg.reset_low, self._cfg.reset_high = low, high
        if mask is not None and self.copy and self._has_reset:
            out["obs"].copy_(self._last_obs)  # unmasked lanes keep their previous observation (:214-246)
        _lib.check(
            self._lib.b2e_cartpole_reset(C.byref(self._batch), C.byref(self._cfg),
                                         ptr(None if mask is None else mask.view(torch.uint8)), ptr(self._state),
                                         ptr(self._ctrl), ptr(self._rng), ptr(out["obs"]), self._stream),
            "b2e_cartpole_reset",
        )
        self._cfg.reset_low, self._cfg.reset_high = -0.05, 0.05  # autoreset uses the defaults (env.reset())
        self._last_obs = out["obs"]

    def _step_kernel(self, actions, out):
        _lib.check(
            self._lib.b2e_cartpole_step(C.byref(self._batch), C.byref(self._cfg), ptr(actions), ptr(self._state),
                                        ptr(self._ctrl), ptr(self._rng), ptr(out["obs"]), ptr(out["reward"]),
                                        ptr(out["terminated"]), ptr(out["truncated"]), ptr(out.get("final_obs")),
                                        self._stream),
            "b2e_cartpole_step",
        )
        self._last_obs = out["obs"]

    def _step_info(self, out):
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            done = out["terminated"] | out["truncated"]
            # dense variant of SyncVectorEnv's object array (sync_vector_env.py:311-317): rows where _final_obs is set
            return {"final_obs": out["final_obs"], "_final_obs": done, "final_info": {}, "_final_info": done}
        return {}

    # -- fused multi-step path -----------------------------------------------------------------------------------
    def rollout(self, num_steps: int, actions=None, return_actions: bool = False):
        """``num_steps`` fused step()+autoreset calls in ONE launch; state stays in registers, the
        ``[num_steps, N, ...]`` trajectory is streamed to HBM.

        ``actions``: ``(num_steps, N)`` integer tensor/array, or None for uniform random actions drawn on the device
        (Philox4x32-10 keyed by seed / env index / call counter -- the synthetic random-action workload of
        ``gymnasium.utils.performance.benchmark_vector_step``, performance.py:57-103).
        Returns dict(obs float32 [K,N,4], reward float32 [K,N], terminated/truncated bool [K,N][, actions uint8 [K,N]]).
        """
        if self.autoreset_mode != AutoresetMode.NEXT_STEP:
            raise ValueError("rollout() supports AutoresetMode.NEXT_STEP only")
        if not self._has_reset:
            from .. import errors

            raise errors.ResetNeeded("Cannot call env.rollout() before calling env.reset()")
        K, n, dev = int(num_steps), self.num_envs, self.device
        with torch.cuda.device(dev):
            act = None
            if actions is not None:
                act = torch.as_tensor(actions)
                if tuple(act.shape) != (K, n):
                    raise ValueError(f"actions must have shape ({K}, {n}), got {tuple(act.shape)}")
                if act.dtype not in (torch.int64, torch.int32, torch.uint8):
                    act = act.to(torch.int64)
                act = act.to(dev).contiguous()
                self._batch.action_dtype = {torch.int64: 0, torch.int32: 1, torch.uint8: 2}[act.dtype]
            out = {
                "obs": torch.empty((K, n, 4), dtype=torch.float32, device=dev),
                "reward": torch.empty((K, n), dtype=torch.float32, device=dev),
                "terminated": torch.empty((K, n), dtype=torch.bool, device=dev),
                "truncated": torch.empty((K, n), dtype=torch.bool, device=dev),
            }
            if actions is None and return_actions:
                out["actions"] = torch.empty((K, n), dtype=torch.uint8, device=dev)
            _lib.check(
                self._lib.b2e_cartpole_rollout(C.byref(self._batch), C.byref(self._cfg), K, ptr(act),
                                               ptr(out.get("actions")), ptr(self._state), ptr(self._ctrl),
                                               ptr(self._rng), ptr(out["obs"]), ptr(out["reward"]),
                                               ptr(out["terminated"]), ptr(out["truncated"]), self._stream),
                "b2e_cartpole_rollout",
            )
            self._batch.call_counter += K
        return out

"""LunarLander-v3 on the B200 engine.

Mirrors ``LunarLander`` (gymnasium/envs/box2d/lunar_lander.py:79-665) behind the vector API with SyncVectorEnv's
conventions; the rigid-body solve that the reference delegates to the Box2D wheel runs in
``gymnasium_b200/csrc/lunarlander.cu``.  Discrete actions, ``enable_wind=False`` (the registered defaults).
Numeric parity with the real Box2D wheel is unpinned (it cannot be installed here); see DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _lib
from .._api import AutoresetMode, Box, Discrete
from ..vector_env import B200VectorEnv, ptr


class LunarLanderVectorEnv(B200VectorEnv):
    """N LunarLander-v3 envs.  Observation ``(N, 8) float32``, action ``(N,) int64`` in {0..3}, reward float64."""

    metadata = {"render_modes": [], "render_fps": 50, "autoreset_mode": AutoresetMode.NEXT_STEP}

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 1000, continuous: bool = False,
                 gravity: float = -10.0, enable_wind: bool = False, wind_power: float = 15.0,
                 turbulence_power: float = 1.5, render_mode: str | None = None, **engine_kwargs):
        if continuous:
            raise NotImplementedError("continuous=True is not implemented by gymnasium_b200 (discrete LunarLander-v3 only)")
        if enable_wind:
            raise NotImplementedError("enable_wind=True is not implemented by gymnasium_b200")
        assert -12.0 < gravity and gravity < 0.0, f"gravity (current value: {gravity}) must be between -12 and 0"  # :233
        low = np.array([-2.5, -2.5, -10.0, -10.0, -2 * math.pi, -10.0, -0.0, -0.0]).astype(np.float32)  # :258-293
        high = np.array([2.5, 2.5, 10.0, 10.0, 2 * math.pi, 10.0, 1.0, 1.0]).astype(np.float32)
        super().__init__(num_envs, Box(low, high), Discrete(4), max_episode_steps=max_episode_steps,
                         render_mode=render_mode, **engine_kwargs)
        self.gravity = float(gravity)
        n, dev = self.num_envs, self.device
        self._cfg = _lib.LunarLanderCfg(gravity=self.gravity, enable_wind=0, continuous=0)
        words = self._lib.b2e_lunarlander_state_words()
        self._s = {
            "bodies": torch.zeros((21, n), dtype=torch.float32, device=dev),
            "joints": torch.zeros((8, n), dtype=torch.float32, device=dev),
            "terrain": torch.zeros((11, n), dtype=torch.float32, device=dev),
            "fat": torch.zeros((12, n), dtype=torch.float32, device=dev),
            "contacts": torch.zeros((words, n), dtype=torch.int32, device=dev),
            "flags": torch.zeros(n, dtype=torch.int32, device=dev),
            "prev_shaping": torch.zeros(n, dtype=torch.float64, device=dev),
            "work": torch.zeros(n, dtype=torch.int32, device=dev),    # scheduling key of every env after its last step
            "order": torch.zeros(9 * n + 64, dtype=torch.int32, device=dev),  # scratch: thread slot -> env of a grouped launch
        }
        self._state = _lib.LunarLanderState(ctrl=self._ctrl.data_ptr(), rng=ptr(self._rng),
                                            **{k: v.data_ptr() for k, v in self._s.items()})

    def _alloc_outputs(self):
        n = self.num_envs
        layout = {"obs": ((n, 8), torch.float32), "reward": ((n,), torch.float64),
                  "terminated": ((n,), torch.bool), "truncated": ((n,), torch.bool)}
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            layout["final_obs"] = ((n, 8), torch.float32)
        out = self._alloc_packed(layout)
        if "final_obs" in out:
            out["final_obs"].zero_()
        return out

    def _reset_kernel(self, mask, options, out):
        if mask is not None and self.copy and self._has_reset:
            out["obs"].copy_(self._last_obs)
        _lib.check(
            self._lib.b2e_lunarlander_reset(C.byref(self._batch), C.byref(self._cfg), C.byref(self._state),
                                            ptr(None if mask is None else mask.view(torch.uint8)), ptr(out["obs"]),
                                            self._stream),
            "b2e_lunarlander_reset",
        )
        self._last_obs = out["obs"]

    def _step_kernel(self, actions, out):
        _lib.check(
            self._lib.b2e_lunarlander_step(C.byref(self._batch), C.byref(self._cfg), C.byref(self._state), ptr(actions),
                                           ptr(out["obs"]), ptr(out["reward"]), ptr(out["terminated"]),
                                           ptr(out["truncated"]), ptr(out.get("final_obs")), self._stream),
            "b2e_lunarlander_step",
        )
        self._last_obs = out["obs"]

    def _step_info(self, out):
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            done = out["terminated"] | out["truncated"]
            return {"final_obs": out["final_obs"], "_final_obs": done, "final_info": {}, "_final_info": done}
        return {}

    # introspection used by the parity tests ------------------------------------------------------------------------
    def body_state(self) -> torch.Tensor:
        """float32 ``(N, 3, 7)``: per body (lander, legs[0], legs[1]) c.x, c.y, angle, v.x, v.y, w, sleepTime."""
        return self._s["bodies"].view(3, 7, self.num_envs).permute(2, 0, 1).contiguous()

    def contact_overflow(self) -> bool:
        """True if any env ever ran out of contact slots (never expected for this scene)."""
        return bool(((self._s["flags"] >> 12) & 1).any())

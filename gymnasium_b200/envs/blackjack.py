"""Blackjack-v1 on the B200 engine.

Mirrors ``BlackjackEnv`` (gymnasium/envs/toy_text/blackjack.py:163-238) behind the vector API with SyncVectorEnv's
conventions; the card game itself runs in ``gymnasium_b200/csrc/blackjack.cu``.  The reference's observation space is
``Tuple(Discrete(32), Discrete(11), Discrete(2))``; SyncVectorEnv batches it into a tuple of three ``(N,)`` int64 arrays
(gymnasium/vector/utils/space_utils.py:120-131), and so does this class: the kernel writes one ``(3, N)`` int64 buffer
and the three rows are returned as the tuple (no copies).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._api import AutoresetMode, Discrete, Tuple
from ..vector_env import B200VectorEnv, ptr


class BlackjackVectorEnv(B200VectorEnv):
    """N Blackjack-v1 tables.  Observation: tuple ``(player sum, dealer's showing card, usable ace)`` of ``(N,)`` int64;
    actions 0 = stick, 1 = hit; reward float64 in {-1, 0, 1, 1.5}; ``info = {}``."""

    metadata = {"render_modes": [], "render_fps": 4, "autoreset_mode": AutoresetMode.NEXT_STEP}

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = None, natural: bool = False, sab: bool = False,
                 render_mode: str | None = None, **engine_kwargs):
        super().__init__(num_envs, Tuple((Discrete(32), Discrete(11), Discrete(2))), Discrete(2),
                         max_episode_steps=max_episode_steps, render_mode=render_mode, **engine_kwargs)
        n, dev = self.num_envs, self.device
        self.natural, self.sab = bool(natural), bool(sab)
        self._hand = torch.zeros(n, dtype=torch.int32, device=dev)
        self._u32buf = torch.zeros(n, dtype=torch.int64, device=dev)  # PCG64's 32-bit word buffer (numpy: has_uint32/uinteger)
        self._cfg = _lib.BlackjackCfg(natural=int(self.natural), sab=int(self.sab), hand=self._hand.data_ptr(),
                                      u32buf=self._u32buf.data_ptr(), ctrl=self._ctrl.data_ptr(), rng=ptr(self._rng))

    def _on_streams_seeded(self, lanes):
        """A freshly seeded numpy Generator starts with an empty 32-bit buffer."""
        if lanes is None:
            self._u32buf.zero_()
        else:
            self._u32buf.masked_fill_(lanes, 0)

    def _alloc_outputs(self):
        n = self.num_envs
        layout = {"obs": ((3, n), torch.int64), "reward": ((n,), torch.float64), "terminated": ((n,), torch.bool),
                  "truncated": ((n,), torch.bool)}
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            layout["final_obs"] = ((3, n), torch.int64)
        out = self._alloc_packed(layout)
        if "final_obs" in out:
            out["final_obs"].zero_()
        return out

    def _reset_kernel(self, mask, options, out):
        if mask is not None and self._has_reset and self.copy:
            out["obs"].copy_(self._last_obs)  # unmasked tables keep their previous observation
        _lib.check(
            self._lib.b2e_blackjack_reset(C.byref(self._batch), C.byref(self._cfg),
                                          ptr(None if mask is None else mask.view(torch.uint8)), ptr(out["obs"]), self._stream),
            "b2e_blackjack_reset",
        )
        self._last_obs = out["obs"]

    def _step_kernel(self, actions, out):
        _lib.check(
            self._lib.b2e_blackjack_step(C.byref(self._batch), C.byref(self._cfg), ptr(actions), ptr(out["obs"]),
                                         ptr(out["reward"]), ptr(out["terminated"]), ptr(out["truncated"]),
                                         ptr(out.get("final_obs")), self._stream),
            "b2e_blackjack_step",
        )
        self._last_obs = out["obs"]

    @staticmethod
    def _as_tuple(obs):
        return (obs[0], obs[1], obs[2])

    def _reset_info(self, out, mask):
        return {}

    def _step_info(self, out):
        if self.autoreset_mode != AutoresetMode.SAME_STEP:
            return {}
        done = out["terminated"] | out["truncated"]
        return {"final_obs": self._as_tuple(out["final_obs"]), "_final_obs": done, "final_info": {}, "_final_info": done}

    def _host_obs(self, host):
        return self._as_tuple(host["obs"])

    def reset(self, *, seed=None, options=None):
        obs, info = super().reset(seed=seed, options=options)
        return self._as_tuple(obs), info

    def step(self, actions):
        obs, reward, terminated, truncated, info = super().step(actions)
        return self._as_tuple(obs), reward, terminated, truncated, info

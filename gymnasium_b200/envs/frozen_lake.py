"""FrozenLake-v1 on the B200 engine.

Mirrors ``FrozenLakeEnv`` (gymnasium/envs/toy_text/frozen_lake.py:84-348) behind the vector API with SyncVectorEnv's
conventions.  The host builds the transition table from the map the way ``FrozenLakeEnv.__init__`` does (:256-300) and
packs it for ``gymnasium_b200/csrc/frozenlake.cu``; integer state, rewards and flags are bit-exact with the reference
for identical seeds.  Deviation: ``info["prob"]`` is always float64 (1.0 on reset calls).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._api import AutoresetMode, Discrete
from ..vector_env import B200VectorEnv, ptr

LEFT, DOWN, RIGHT, UP = 0, 1, 2, 3

# frozen_lake.py:14-31 (environment data)
MAPS = {
    "4x4": ["SFFF", "FHFH", "FFFH", "HFFG"],
    "8x8": [
        "SFFFFFFF",
        "FFFFFFFF",
        "FFFHFFFF",
        "FFFFFHFF",
        "FFFHFFFF",
        "FHHFFFHF",
        "FHFFHFHF",
        "FFFHFFFG",
    ],
}


def pack_transition_table(desc, is_slippery: bool, success_rate: float, reward_schedule):
    """Pack P[s][a] (frozen_lake.py:256-300) for the kernel.

    Returns (table uint32 [nS*4*3], cum3, p3, isd_cum float64 [nS], nS).  Entry layout (include/b200env.h):
    next_state | done << 16 | reward_class << 17 | n_out << 20, reward_class indexing (G, H, F).
    """
    rows = [r.decode() if isinstance(r, bytes) else str(r) for r in desc]
    nrow, ncol = len(rows), len(rows[0])
    if any(len(r) != ncol for r in rows):
        raise ValueError("map rows must have equal length")
    nS, nA = nrow * ncol, 4
    if nS > 65536:
        raise ValueError("maps above 65536 tiles are not supported")
    fail_rate = (1.0 - success_rate) / 2.0  # :258

    def move(row, col, a):  # :263-272
        if a == LEFT:
            col = max(col - 1, 0)
        elif a == DOWN:
            row = min(row + 1, nrow - 1)
        elif a == RIGHT:
            col = min(col + 1, ncol - 1)
        elif a == UP:
            row = max(row - 1, 0)
        return row, col

    def outcome(row, col, a, n_out):  # :274-283
        r2, c2 = move(row, col, a)
        letter = rows[r2][c2]
        done = letter in "GH"
        rclass = "GHF".index(letter if letter in "GHF" else "F")
        return (r2 * ncol + c2) | (int(done) << 16) | (rclass << 17) | (n_out << 20)

    table = np.zeros((nS, nA, 3), dtype=np.uint32)
    for row in range(nrow):
        for col in range(ncol):
            s = row * ncol + col
            for a in range(nA):
                if rows[row][col] in "GH":  # :290-291 -> (1.0, s, 0, True); reward 0 is literal, not the schedule
                    table[s, a, :] = s | (1 << 16) | (3 << 17) | (1 << 20)
                elif is_slippery:  # :293-300
                    for k, b in enumerate([(a - 1) % 4, a, (a + 1) % 4]):
                        table[s, a, k] = outcome(row, col, b, 3)
                else:  # :301-302
                    table[s, a, :] = outcome(row, col, a, 1)
    p3 = np.array([fail_rate, success_rate, fail_rate], dtype=np.float64)
    cum3 = np.cumsum(p3)  # toy_text/utils.py:7
    isd = np.array([c == "S" for r in rows for c in r], dtype=np.float64)  # :249-250
    isd /= isd.sum()
    return table.reshape(-1), cum3, p3, np.cumsum(isd), nS, (nrow, ncol)


class TabularVectorEnv(B200VectorEnv):
    """N copies of a tabular MDP ``P[s][a] = [(p, s', r, done), ...]`` with 1 or 3 outcomes per (s, a) -- the structure
    every ``gymnasium/envs/toy_text`` grid world has -- stepped by ``csrc/frozenlake.cu``.  Observation ``(N,) int64``,
    reward float64, ``info = {"prob": float64 (N,), "_prob": bool (N,)}`` as ``SyncVectorEnv`` batches it."""

    metadata = {"render_modes": [], "render_fps": 4, "autoreset_mode": AutoresetMode.NEXT_STEP}

    def __init__(self, num_envs, n_states, n_actions, table, cum3, p3, isd_cum, rewards, *, max_episode_steps,
                 render_mode=None, **engine_kwargs):
        super().__init__(num_envs, Discrete(int(n_states)), Discrete(int(n_actions)),
                         max_episode_steps=max_episode_steps, render_mode=render_mode, **engine_kwargs)
        dev = self.device
        self._table = torch.from_numpy(np.ascontiguousarray(table, dtype=np.uint32).view(np.int32)).to(dev)
        self._isd_cum = torch.from_numpy(np.ascontiguousarray(isd_cum, dtype=np.float64)).to(dev)
        self._cfg = _lib.FrozenLakeCfg(n_states=int(n_states), n_actions=int(n_actions), table=self._table.data_ptr(),
                                       isd_cum=self._isd_cum.data_ptr())
        for k in range(3):
            self._cfg.cum3[k], self._cfg.p3[k] = float(cum3[k]), float(p3[k])
            self._cfg.rewards[k] = float(rewards[k])
        self._pstate = torch.zeros(self.num_envs, dtype=torch.int32, device=dev)

    def _alloc_outputs(self):
        n = self.num_envs
        layout = {"obs": ((n,), torch.int64), "reward": ((n,), torch.float64), "prob": ((n,), torch.float64),
                  "terminated": ((n,), torch.bool), "truncated": ((n,), torch.bool)}
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            layout["final_obs"] = ((n,), torch.int64)
            layout["final_prob"] = ((n,), torch.float64)
        out = self._alloc_packed(layout)
        if "final_obs" in out:
            out["final_obs"].zero_()
            out["final_prob"].zero_()
        return out

    def _reset_kernel(self, mask, options, out):
        if mask is not None:
            if self._has_reset and self.copy:
                out["obs"].copy_(self._last_obs)
            out["prob"].zero_()
        _lib.check(
            self._lib.b2e_frozenlake_reset(C.byref(self._batch), C.byref(self._cfg),
                                           ptr(None if mask is None else mask.view(torch.uint8)), ptr(self._pstate),
                                           ptr(self._ctrl), ptr(self._rng), ptr(out["obs"]), ptr(out["prob"]),
                                           self._stream),
            "b2e_frozenlake_reset",
        )
        self._last_obs = out["obs"]

    def _reset_info(self, out, mask):
        if isinstance(out["prob"], np.ndarray):
            m = np.ones(self.num_envs, dtype=np.bool_) if mask is None else mask.copy()
        else:
            m = torch.ones(self.num_envs, dtype=torch.bool, device=self.device) if mask is None else mask.clone()
        return {"prob": out["prob"], "_prob": m}

    def _step_kernel(self, actions, out):
        _lib.check(
            self._lib.b2e_frozenlake_step(C.byref(self._batch), C.byref(self._cfg), ptr(actions), ptr(self._pstate),
                                          ptr(self._ctrl), ptr(self._rng), ptr(out["obs"]), ptr(out["reward"]),
                                          ptr(out["terminated"]), ptr(out["truncated"]), ptr(out["prob"]),
                                          ptr(out.get("final_obs")), ptr(out.get("final_prob")), self._stream),
            "b2e_frozenlake_step",
        )
        self._last_obs = out["obs"]

    def _step_info(self, out):
        host = isinstance(out["prob"], np.ndarray)
        if host:
            all_true = np.ones(self.num_envs, dtype=np.bool_)
        else:
            if not hasattr(self, "_all_true"):
                self._all_true = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
            all_true = self._all_true
        info = {"prob": out["prob"], "_prob": all_true}
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            done = out["terminated"] | out["truncated"]
            info.update({"final_obs": out["final_obs"], "_final_obs": done,
                         "final_info": {"prob": out["final_prob"], "_prob": done}, "_final_info": done})
        return info

    def rollout(self, num_steps: int, actions=None, return_actions: bool = False):
        """``num_steps`` fused step()+autoreset calls in one launch (see CartPoleVectorEnv.rollout).
        Returns dict(obs int64 [K,N], reward float32 [K,N], terminated/truncated bool [K,N][, actions uint8 [K,N]])."""
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
                "obs": torch.empty((K, n), dtype=torch.int64, device=dev),
                "reward": torch.empty((K, n), dtype=torch.float32, device=dev),
                "terminated": torch.empty((K, n), dtype=torch.bool, device=dev),
                "truncated": torch.empty((K, n), dtype=torch.bool, device=dev),
            }
            if actions is None and return_actions:
                out["actions"] = torch.empty((K, n), dtype=torch.uint8, device=dev)
            _lib.check(
                self._lib.b2e_frozenlake_rollout(C.byref(self._batch), C.byref(self._cfg), K, ptr(act),
                                                 ptr(out.get("actions")), ptr(self._pstate), ptr(self._ctrl),
                                                 ptr(self._rng), ptr(out["obs"]), ptr(out["reward"]),
                                                 ptr(out["terminated"]), ptr(out["truncated"]), self._stream),
                "b2e_frozenlake_rollout",
            )
            self._batch.call_counter += K
        return out


class FrozenLakeVectorEnv(TabularVectorEnv):
    """N FrozenLake-v1 envs.  Observation ``(N,) int64``, action ``(N,) int64`` in {0..3}, reward float64, and
    ``info = {"prob": float64 (N,), "_prob": bool (N,)}`` as ``SyncVectorEnv`` batches it."""

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 100, desc=None, map_name: str | None = "4x4",
                 is_slippery: bool = True, success_rate: float = 1.0 / 3.0, reward_schedule=(1, 0, 0),
                 render_mode: str | None = None, **engine_kwargs):
        if desc is None and map_name is None:
            raise NotImplementedError("random maps (desc=None, map_name=None) are not supported; pass desc=")
        if desc is None:
            desc = MAPS[map_name]
        table, cum3, p3, isd_cum, nS, shape = pack_transition_table(desc, is_slippery, success_rate, reward_schedule)
        super().__init__(num_envs, nS, 4, table, cum3, p3, isd_cum, reward_schedule,
                         max_episode_steps=max_episode_steps, render_mode=render_mode, **engine_kwargs)
        self.desc = np.asarray([list(r) for r in desc], dtype="c")
        self.nrow, self.ncol = shape

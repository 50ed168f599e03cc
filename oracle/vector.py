"""Oracle: SyncVectorEnv batching / NEXT_STEP autoreset / TimeLimit semantics, vectorised with numpy.

Oracle only (see oracle/__init__.py).  Restates
  * ``SyncVectorEnv.reset``  gymnasium/vector/sync_vector_env.py:187-264  (seed -> seed+i :205-208,
    reset_mask validation :214-231)
  * ``SyncVectorEnv.step``   gymnasium/vector/sync_vector_env.py:266-337  (NEXT_STEP :279-292,
    SAME_STEP :302-319, DISABLED :293-301, ``_autoreset_envs = term | trunc`` :329, reward float64 :171)
  * ``TimeLimit.step/reset`` gymnasium/wrappers/common.py:116-151 (``elapsed += 1``;
    ``truncated = elapsed >= max_episode_steps``; reset -> 0)
  * per-env RNG  gymnasium/utils/seeding.py:10-42, ``Env.reset(seed)`` gymnasium/core.py:116-159
    (no seed => the stream continues).
Each sub-env owns ``numpy.random.Generator(PCG64(SeedSequence(seed+i)))`` exactly as the reference does
(numpy is the reference's own RNG dependency; oracle/np_rng.py restates it and is pinned against it).
"""
from __future__ import annotations

import numpy as np

NEXT_STEP, SAME_STEP, DISABLED = "NextStep", "SameStep", "Disabled"


class OracleVectorEnv:
    """Family-independent part.  Subclasses implement ``_reset_env(i, options)`` and ``_step_env(i, action)``
    or the batched ``_step_batch(actions, lanes)`` fast path."""

    def __init__(self, num_envs: int, max_episode_steps: int | None, autoreset_mode: str = NEXT_STEP):
        self.num_envs = int(num_envs)
        self.max_episode_steps = max_episode_steps
        self.autoreset_mode = autoreset_mode
        self.rngs: list[np.random.Generator | None] = [None] * self.num_envs
        self.elapsed = np.zeros(self.num_envs, dtype=np.int64)
        self.autoreset = np.zeros(self.num_envs, dtype=bool)

    # -- hooks -------------------------------------------------------------------------------------
    def _reset_env(self, i: int, options) -> dict:
        raise NotImplementedError

    def _step_lanes(self, lanes: np.ndarray, actions: np.ndarray):
        """Step the given lanes; returns (reward f64[len], terminated bool[len], info dict of arrays)."""
        raise NotImplementedError

    def _obs(self) -> np.ndarray:
        raise NotImplementedError

    def _reset_info(self, lanes) -> dict:
        return {}

    # -- API ---------------------------------------------------------------------------------------
    def _rng(self, i: int) -> np.random.Generator:
        if self.rngs[i] is None:  # core.py:226-235 lazy self-seeding
            self.rngs[i] = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
        return self.rngs[i]

    def reset(self, *, seed=None, options=None):
        n = self.num_envs
        if seed is None:
            seeds = [None] * n
        elif isinstance(seed, (int, np.integer)):
            seeds = [int(seed) + i for i in range(n)]
        else:
            seeds = list(seed)
        if len(seeds) != n:
            raise ValueError(
                f"If seeds are passed as a list the length must match num_envs={n} but got length={len(seeds)}."
            )
        mask = np.ones(n, dtype=bool)
        if options is not None and "reset_mask" in options:
            options = dict(options)
            mask = options.pop("reset_mask")
            if not isinstance(mask, np.ndarray):
                raise TypeError(f"`options['reset_mask']` must be a numpy array, got {type(mask)}")
            if mask.shape != (n,):
                raise ValueError(f"`options['reset_mask']` must have shape `({n},)`, got {mask.shape}")
            if mask.dtype != np.bool_:
                raise TypeError(f"`options['reset_mask']` must have `dtype=np.bool_`, got {mask.dtype}")
            if not np.any(mask):
                raise ValueError(
                    f"`options['reset_mask']` must contain a boolean array with at least one True value, got reset_mask={mask}"
                )
        lanes = np.flatnonzero(mask)
        for i in lanes:
            if seeds[i] is not None:
                self.rngs[i] = np.random.Generator(np.random.PCG64(np.random.SeedSequence(int(seeds[i]))))
            self._reset_env(int(i), options)
        self.elapsed[lanes] = 0
        self.autoreset[lanes] = False
        info = {}
        for k, v in self._reset_info(lanes).items():
            full = np.zeros_like(v)
            full[lanes] = v[lanes]
            info[k], info["_" + k] = full, mask.copy()
        return self._obs().copy(), info

    def step(self, actions):
        n = self.num_envs
        actions = np.asarray(actions)
        if actions.shape[0] != n:
            raise ValueError(f"expected {n} actions, got {actions.shape[0]}")
        reward = np.zeros(n, dtype=np.float64)
        term = np.zeros(n, dtype=bool)
        trunc = np.zeros(n, dtype=bool)
        info: dict = {}
        if self.autoreset_mode == NEXT_STEP:
            reset_lanes = np.flatnonzero(self.autoreset)
            step_lanes = np.flatnonzero(~self.autoreset)
        else:
            if self.autoreset_mode == DISABLED:
                assert not self.autoreset.any(), f"{self.autoreset=}"
            reset_lanes = np.zeros(0, dtype=np.int64)
            step_lanes = np.arange(n)
        for i in reset_lanes:  # sync_vector_env.py:279-284 (reward 0, flags False, action ignored)
            self._reset_env(int(i), None)
        self.elapsed[reset_lanes] = 0
        if len(step_lanes):
            r, t, sinfo = self._step_lanes(step_lanes, actions[step_lanes])
            reward[step_lanes] = r
            term[step_lanes] = t
            self.elapsed[step_lanes] += 1
            if self.max_episode_steps is not None:
                trunc[step_lanes] = self.elapsed[step_lanes] >= self.max_episode_steps
            for k, v in sinfo.items():
                full = np.zeros((n,) + v.shape[1:], dtype=v.dtype)
                full[step_lanes] = v
                m = np.zeros(n, dtype=bool)
                m[step_lanes] = True
                info[k], info["_" + k] = full, m
        rinfo = self._reset_info(reset_lanes) if len(reset_lanes) else {}
        for k, v in rinfo.items():
            if k.startswith("_"):
                continue
            if k not in info:
                info[k] = np.zeros((n,) + v.shape[1:], dtype=v.dtype)
                info["_" + k] = np.zeros(n, dtype=bool)
            info[k][reset_lanes] = v[reset_lanes]
            info["_" + k][reset_lanes] = True
        obs = self._obs().copy()
        done = term | trunc
        if self.autoreset_mode == SAME_STEP and done.any():
            lanes = np.flatnonzero(done)
            info["final_obs"] = np.full(n, None, dtype=object)
            for i in lanes:
                info["final_obs"][i] = obs[i].copy()
                self._reset_env(int(i), None)
            info["_final_obs"] = done.copy()
            # sync_vector_env.py:311-319: the step's info moves to final_info, the reset's info takes its place
            final_info = {}
            for k in [k for k in info if not k.startswith("_") and k != "final_obs"]:
                fv = np.zeros_like(info[k]); fv[lanes] = info[k][lanes]
                final_info[k], final_info["_" + k] = fv, done.copy()
            if final_info:
                info["final_info"], info["_final_info"] = final_info, done.copy()
            for k, v in self._reset_info(lanes).items():
                info[k][lanes] = v[lanes]
            self.elapsed[lanes] = 0
            obs = self._obs().copy()
            self.autoreset[:] = False
        else:
            self.autoreset = done if self.autoreset_mode != SAME_STEP else np.zeros(n, dtype=bool)
        return obs, reward, term, trunc, info

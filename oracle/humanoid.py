"""ctypes front end of oracle/humanoid.c (the MuJoCo-subset Humanoid-v5 oracle).  Oracle only; PARITY UNPINNED
(mujoco is absent from this image) -- see the header of humanoid.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libhumanoid_oracle.so")
_lib = None
INFO_KEYS = ["x_position", "y_position", "tendon_length0", "tendon_length1", "tendon_velocity0", "tendon_velocity1",
             "distance_from_origin", "x_velocity", "y_velocity", "reward_survive", "reward_forward", "reward_ctrl",
             "reward_contact"]


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "humanoid.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        l = C.CDLL(_LIB)
        l.hm_create.restype = C.c_void_p
        l.hm_create.argtypes = [C.c_int, C.c_int, C.c_double]
        l.hm_destroy.argtypes = [C.c_void_p]
        l.hm_reset.argtypes = [C.c_void_p] * 5
        l.hm_step.argtypes = [C.c_void_p] * 7
        l.hm_model_info.argtypes = [C.c_void_p] * 4
        l.hm_debug.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        l.hm_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


class OracleHumanoid:
    """SyncVectorEnv(Humanoid-v5 x N) semantics: seed+i PCG64 streams, NEXT_STEP autoreset, TimeLimit 1000."""

    def __init__(self, num_envs, max_episode_steps=1000, reset_noise_scale=1e-2):
        self.num_envs = n = int(num_envs)
        self._h = lib().hm_create(n, int(max_episode_steps or 0), float(reset_noise_scale))
        self._obs = np.zeros((n, 348), dtype=np.float64)
        self._info = np.zeros((n, 13), dtype=np.float64)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hm_destroy(self._h)
            self._h = None

    def _info_dict(self):
        return {k: self._info[:, i].copy() for i, k in enumerate(INFO_KEYS)}

    def reset(self, *, seed=None, options=None):
        n = self.num_envs
        seeds = None
        if seed is not None:
            seeds = np.array([seed + i for i in range(n)] if isinstance(seed, (int, np.integer)) else list(seed),
                             dtype=np.uint64)
        mask = None
        if options is not None and "reset_mask" in options:
            mask = np.ascontiguousarray(options["reset_mask"]).astype(np.uint8)
        lib().hm_reset(self._h, None if seeds is None else seeds.ctypes.data, None if mask is None else mask.ctypes.data,
                       self._obs.ctypes.data, self._info.ctypes.data)
        return self._obs.copy(), self._info_dict()

    def step(self, actions):
        n = self.num_envs
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (n, 17):
            raise ValueError(f"Action dimension mismatch. Expected {(n, 17)}, found {a.shape}")
        reward = np.zeros(n, dtype=np.float64)
        term = np.zeros(n, dtype=np.uint8)
        trunc = np.zeros(n, dtype=np.uint8)
        lib().hm_step(self._h, a.ctypes.data, self._obs.ctypes.data, reward.ctypes.data, term.ctypes.data,
                      trunc.ctypes.data, self._info.ctypes.data)
        return self._obs.copy(), reward, term.astype(bool), trunc.astype(bool), self._info_dict()

    def step_inplace(self, actions_f32):
        """step() without per-call allocations or copies (results stay in internal buffers): the call the multi-threaded CPU
        baseline of bench.py makes, so that almost no time is spent holding the GIL.  `actions_f32`: contiguous float32 [n, 17]."""
        if not hasattr(self, "_scratch"):
            n = self.num_envs
            self._scratch = (np.zeros(n, dtype=np.float64), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8))
            r, te, tr = self._scratch
            self._ptrs = (self._obs.ctypes.data, r.ctypes.data, te.ctypes.data, tr.ctypes.data, self._info.ctypes.data)
        lib().hm_step(self._h, actions_f32.ctypes.data, *self._ptrs)

    def model_info(self):
        mass = np.zeros(14); misc = np.zeros(8); inv = np.zeros(14 * 2 + 23)
        lib().hm_model_info(self._h, mass.ctypes.data, misc.ctypes.data, inv.ctypes.data)
        return mass, misc, inv

    def debug(self, i=0):
        qpos = np.zeros(24); qvel = np.zeros(23); qacc = np.zeros(23); counts = np.zeros(3, dtype=np.int32)
        lib().hm_debug(self._h, i, qpos.ctypes.data, qvel.ctypes.data, qacc.ctypes.data, counts.ctypes.data)
        return qpos, qvel, qacc, counts

    def set_state(self, i, qpos, qvel):
        qpos = np.ascontiguousarray(qpos, dtype=np.float64); qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        lib().hm_set_state(self._h, i, qpos.ctypes.data, qvel.ctypes.data)

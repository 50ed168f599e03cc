"""ctypes front end of oracle/hopper.c (the MuJoCo-subset Hopper-v5 oracle).  Oracle only; PARITY UNPINNED (mujoco is absent
from this image) -- see the header of hopper.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libhopper_oracle.so")
_lib = None
INFO_KEYS = ["x_position", "z_distance_from_origin", "x_velocity", "reward_forward", "reward_ctrl", "reward_survive"]
NB, NQ, NV, NU, OBS = 5, 6, 6, 3, 11


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hopper.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        l = C.CDLL(_LIB)
        l.hp_create.restype = C.c_void_p
        l.hp_create.argtypes = [C.c_int, C.c_int, C.c_double]
        l.hp_destroy.argtypes = [C.c_void_p]
        l.hp_reset.argtypes = [C.c_void_p] * 5
        l.hp_step.argtypes = [C.c_void_p] * 7
        l.hp_model_info.argtypes = [C.c_void_p] * 4
        l.hp_debug.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        l.hp_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


class OracleHopper:
    """SyncVectorEnv(Hopper-v5 x N) semantics: seed+i PCG64 streams, NEXT_STEP autoreset, TimeLimit 1000."""

    def __init__(self, num_envs, max_episode_steps=1000, reset_noise_scale=5e-3):
        self.num_envs = n = int(num_envs)
        self._h = lib().hp_create(n, int(max_episode_steps or 0), float(reset_noise_scale))
        self._obs = np.zeros((n, OBS), dtype=np.float64)
        self._info = np.zeros((n, len(INFO_KEYS)), dtype=np.float64)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hp_destroy(self._h)
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
        lib().hp_reset(self._h, None if seeds is None else seeds.ctypes.data, None if mask is None else mask.ctypes.data,
                       self._obs.ctypes.data, self._info.ctypes.data)
        return self._obs.copy(), self._info_dict()

    def step(self, actions):
        n = self.num_envs
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (n, NU):
            raise ValueError(f"Action dimension mismatch. Expected {(n, NU)}, found {a.shape}")
        reward = np.zeros(n, dtype=np.float64)
        term = np.zeros(n, dtype=np.uint8)
        trunc = np.zeros(n, dtype=np.uint8)
        lib().hp_step(self._h, a.ctypes.data, self._obs.ctypes.data, reward.ctypes.data, term.ctypes.data,
                      trunc.ctypes.data, self._info.ctypes.data)
        return self._obs.copy(), reward, term.astype(bool), trunc.astype(bool), self._info_dict()

    def step_inplace(self, actions_f32):
        if not hasattr(self, "_scratch"):
            n = self.num_envs
            self._scratch = (np.zeros(n, dtype=np.float64), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8))
            r, te, tr = self._scratch
            self._ptrs = (self._obs.ctypes.data, r.ctypes.data, te.ctypes.data, tr.ctypes.data, self._info.ctypes.data)
        lib().hp_step(self._h, actions_f32.ctypes.data, *self._ptrs)

    def model_info(self):
        mass = np.zeros(NB); misc = np.zeros(8); inv = np.zeros(NB * 2 + NV)
        lib().hp_model_info(self._h, mass.ctypes.data, misc.ctypes.data, inv.ctypes.data)
        return mass, misc, inv

    def debug(self, i=0):
        qpos = np.zeros(NQ); qvel = np.zeros(NV); qacc = np.zeros(NV); counts = np.zeros(3, dtype=np.int32)
        xipos = np.zeros((NB, 3))
        lib().hp_debug(self._h, i, qpos.ctypes.data, qvel.ctypes.data, qacc.ctypes.data, counts.ctypes.data, xipos.ctypes.data)
        return qpos, qvel, qacc, counts, xipos

    def set_state(self, i, qpos, qvel):
        qpos = np.ascontiguousarray(qpos, dtype=np.float64); qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        lib().hp_set_state(self._h, i, qpos.ctypes.data, qvel.ctypes.data)

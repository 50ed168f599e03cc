"""ctypes front end of oracle/lunar_lander.c (the Box2D-subset LunarLander-v3 oracle).  Oracle only; PARITY UNPINNED
(Box2D is absent from this image) -- see the header of lunar_lander.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liblunar_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "lunar_lander.c")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        l = C.CDLL(_LIB)
        l.ll_create.restype = C.c_void_p
        l.ll_create.argtypes = [C.c_int, C.c_int, C.c_double]
        l.ll_destroy.argtypes = [C.c_void_p]
        l.ll_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.ll_step.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        l.ll_debug_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        l.ll_terrain.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        l.ll_toi_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        l.ll_configure.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        l.ll_wind_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        l.ll_toi_shortcut_stats.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        l.ll_toi_probe.restype = C.c_int
        l.ll_toi_probe.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
        _lib = l
    return _lib


class OracleLunarLander:
    """SyncVectorEnv(LunarLander-v3 x N) semantics: seed+i PCG64 streams, NEXT_STEP autoreset, TimeLimit 1000."""

    def __init__(self, num_envs, max_episode_steps=1000, gravity=-10.0, continuous=False, enable_wind=False, wind_power=15.0,
                 turbulence_power=1.5):
        self.num_envs = int(num_envs)
        self.continuous = bool(continuous)
        self._h = lib().ll_create(self.num_envs, int(max_episode_steps or 0), float(gravity))
        lib().ll_configure(self._h, int(self.continuous), int(bool(enable_wind)), float(wind_power), float(turbulence_power))
        self._obs = np.zeros((self.num_envs, 8), dtype=np.float32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ll_destroy(self._h)
            self._h = None

    def reset(self, *, seed=None, options=None):
        n = self.num_envs
        seeds = None
        if seed is not None:
            seeds = np.array([seed + i for i in range(n)] if isinstance(seed, (int, np.integer)) else list(seed),
                             dtype=np.uint64)
        mask = None
        if options is not None and "reset_mask" in options:
            mask = np.ascontiguousarray(options["reset_mask"]).astype(np.uint8)
        lib().ll_reset(self._h, None if seeds is None else seeds.ctypes.data, None if mask is None else mask.ctypes.data,
                       self._obs.ctypes.data)
        return self._obs.copy(), {}

    def step(self, actions):
        n = self.num_envs
        if self.continuous:
            a = np.ascontiguousarray(actions, dtype=np.float32)
            assert a.shape == (n, 2)
        else:
            a = np.ascontiguousarray(actions, dtype=np.int64)
            assert a.shape == (n,)
            assert ((a >= 0) & (a < 4)).all(), f"invalid action in {a!r}"
        reward = np.zeros(n, dtype=np.float64)
        term = np.zeros(n, dtype=np.uint8)
        trunc = np.zeros(n, dtype=np.uint8)
        lib().ll_step(self._h, a.ctypes.data, self._obs.ctypes.data, reward.ctypes.data, term.ctypes.data, trunc.ctypes.data)
        return self._obs.copy(), reward, term.astype(bool), trunc.astype(bool), {}

    def step_inplace(self, actions_i64):
        """step() without per-call allocations or copies (results stay in internal buffers): the call the multi-threaded CPU
        baseline of bench.py makes, so that almost no time is spent holding the GIL.  `actions_i64`: contiguous int64 [n]."""
        if not hasattr(self, "_scratch"):
            n = self.num_envs
            self._scratch = (np.zeros(n, dtype=np.float64), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8))
            self._ptrs = tuple(x.ctypes.data for x in (self._obs,) + self._scratch)
        lib().ll_step(self._h, actions_i64.ctypes.data, *self._ptrs)

    def debug_state(self, i=0):
        bodies = np.zeros((3, 7), dtype=np.float32)
        misc = np.zeros(8, dtype=np.float32)
        lib().ll_debug_state(self._h, i, bodies.ctypes.data, misc.ctypes.data)
        return bodies, misc

    def wind_state(self, i=0):
        """(wind_idx, torque_idx) of env i."""
        out = (C.c_int64 * 2)()
        lib().ll_wind_state(self._h, i, out)
        return int(out[0]), int(out[1])

    def toi_stats(self, i=0):
        """(b2TimeOfImpact evaluations, solid TOI events) of env i since its last reset."""
        out = (C.c_long * 2)()
        lib().ll_toi_stats(self._h, i, out)
        return int(out[0]), int(out[1])

    def toi_shortcut_stats(self, i=0):
        """(evaluations the CUDA engine's clearly-separated shortcut covers, of those b2TimeOfImpact did not answer alpha=1)."""
        out = (C.c_long * 2)()
        lib().ll_toi_shortcut_stats(self._h, i, out)
        return int(out[0]), int(out[1])

    def terrain(self, i=0):
        xy = np.zeros((11, 4), dtype=np.float32)
        lib().ll_terrain(self._h, i, xy.ctypes.data)
        return xy


def heuristic(s):
    """The reference's own demo/test policy, gymnasium/envs/box2d/lunar_lander.py:791-842 (discrete branch)."""
    angle_targ = s[0] * 0.5 + s[2] * 1.0
    angle_targ = min(max(angle_targ, -0.4), 0.4)
    hover_targ = 0.55 * np.abs(s[0])
    angle_todo = (angle_targ - s[4]) * 0.5 - (s[5]) * 1.0
    hover_todo = (hover_targ - s[1]) * 0.5 - (s[3]) * 0.5
    if s[6] or s[7]:
        angle_todo = 0
        hover_todo = -(s[3]) * 0.5
    a = 0
    if hover_todo > np.abs(angle_todo) and hover_todo > 0.05:
        a = 2
    elif angle_todo < -0.05:
        a = 3
    elif angle_todo > +0.05:
        a = 1
    return a


def toi_probe(edge, half_extents, c0, a0, c1, a1):
    """The oracle's restated b2TimeOfImpact for a box swept from (c0, a0) to (c1, a1) against an edge -> (state, t)."""
    e = np.asarray(edge, dtype=np.float32).reshape(4)
    p0, p1 = np.asarray(c0, dtype=np.float32), np.asarray(c1, dtype=np.float32)
    t = C.c_float(0)
    st = lib().ll_toi_probe(e.ctypes.data, float(half_extents[0]), float(half_extents[1]), p0.ctypes.data, float(a0),
                            p1.ctypes.data, float(a1), C.byref(t))
    return int(st), float(t.value)

"""ctypes front end of the planar MuJoCo oracle core (oracle/mjc_planar.h: Hopper-v5, Walker2d-v5).  Oracle only; PARITY
UNPINNED (mujoco is absent from this image) -- see the header of mjc_planar.h."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
INFO_KEYS = ["x_position", "z_distance_from_origin", "x_velocity", "reward_forward", "reward_ctrl", "reward_survive"]
# robot -> (library, symbol prefix, nbody, nq = nv, nu)
ROBOTS = {"hopper": ("libhopper_oracle.so", "hp_", 5, 6, 3), "walker2d": ("libwalker2d_oracle.so", "w2_", 8, 9, 6),
          "inverted_pendulum": ("libinverted_pendulum_oracle.so", "ip_", 3, 2, 1),
          "half_cheetah": ("libhalf_cheetah_oracle.so", "hc_", 8, 9, 6),
          # test-only: the inverted pendulum under mj_Euler (checks the integrator HalfCheetah uses against physics)
          "inverted_pendulum_euler": ("libinverted_pendulum_euler_oracle.so", "ipe_", 3, 2, 1)}
OBS_SIZE = {"inverted_pendulum": 4, "inverted_pendulum_euler": 4}  # default: 2 nq - 1 (qpos[1:] | qvel)
_libs = {}


def lib(robot):
    if robot not in _libs:
        name, pre, *_ = ROBOTS[robot]
        path = os.path.join(_HERE, "_build", name)
        srcs = [os.path.join(_HERE, f) for f in (robot + ".c", "mjc_planar.h")]
        if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(f) for f in srcs):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        l = C.CDLL(path)
        f = lambda n: getattr(l, pre + n)  # noqa: E731
        f("create").restype = C.c_void_p
        f("create").argtypes = [C.c_int, C.c_int, C.c_double]
        f("destroy").argtypes = [C.c_void_p]
        f("reset").argtypes = [C.c_void_p] * 5
        f("step").argtypes = [C.c_void_p] * 7
        f("model_info").argtypes = [C.c_void_p] * 4
        f("debug").argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        f("set_state").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        f("contact_forces").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        f("contact_forces").restype = C.c_int
        _libs[robot] = {n: f(n) for n in ("create", "destroy", "reset", "step", "model_info", "debug", "set_state", "contact_forces")}
    return _libs[robot]


class OraclePlanar:
    """SyncVectorEnv(<robot>-v5 x N) semantics: seed+i PCG64 streams, NEXT_STEP autoreset, TimeLimit 1000."""

    robot = "hopper"

    def __init__(self, num_envs, max_episode_steps=1000, reset_noise_scale=5e-3):
        self.num_envs = n = int(num_envs)
        _, _, self.nb, self.nq, self.nu = ROBOTS[self.robot]
        self.nv, self.obs_size = self.nq, OBS_SIZE.get(self.robot, 2 * self.nq - 1)
        self._f = lib(self.robot)
        self._h = self._f["create"](n, int(max_episode_steps or 0), float(reset_noise_scale))
        self._obs = np.zeros((n, self.obs_size), dtype=np.float64)
        self._info = np.zeros((n, len(INFO_KEYS)), dtype=np.float64)

    def __del__(self):
        if getattr(self, "_h", None):
            self._f["destroy"](self._h)
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
        self._f["reset"](self._h, None if seeds is None else seeds.ctypes.data, None if mask is None else mask.ctypes.data,
                         self._obs.ctypes.data, self._info.ctypes.data)
        return self._obs.copy(), self._info_dict()

    def step(self, actions):
        n = self.num_envs
        a = np.ascontiguousarray(actions, dtype=np.float32)
        if a.shape != (n, self.nu):
            raise ValueError(f"Action dimension mismatch. Expected {(n, self.nu)}, found {a.shape}")
        reward = np.zeros(n, dtype=np.float64)
        term = np.zeros(n, dtype=np.uint8)
        trunc = np.zeros(n, dtype=np.uint8)
        self._f["step"](self._h, a.ctypes.data, self._obs.ctypes.data, reward.ctypes.data, term.ctypes.data,
                        trunc.ctypes.data, self._info.ctypes.data)
        return self._obs.copy(), reward, term.astype(bool), trunc.astype(bool), self._info_dict()

    def step_inplace(self, actions_f32):
        """step() without per-call allocations or copies: what the multi-threaded CPU baseline of bench.py calls."""
        if not hasattr(self, "_scratch"):
            n = self.num_envs
            self._scratch = (np.zeros(n, dtype=np.float64), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8))
            r, te, tr = self._scratch
            self._ptrs = (self._obs.ctypes.data, r.ctypes.data, te.ctypes.data, tr.ctypes.data, self._info.ctypes.data)
        self._f["step"](self._h, actions_f32.ctypes.data, *self._ptrs)

    def model_info(self):
        mass = np.zeros(self.nb); misc = np.zeros(8); inv = np.zeros(self.nb * 2 + self.nv)
        self._f["model_info"](self._h, mass.ctypes.data, misc.ctypes.data, inv.ctypes.data)
        return mass, misc, inv

    def debug(self, i=0):
        qpos = np.zeros(self.nq); qvel = np.zeros(self.nv); qacc = np.zeros(self.nv); counts = np.zeros(3, dtype=np.int32)
        xipos = np.zeros((self.nb, 3))
        self._f["debug"](self._h, i, qpos.ctypes.data, qvel.ctypes.data, qacc.ctypes.data, counts.ctypes.data, xipos.ctypes.data)
        return qpos, qvel, qacc, counts, xipos

    def contact_forces(self, i=0):
        """(normal force per contact, contact normals) of env i's last forward evaluation."""
        f = np.zeros(64); nrm = np.zeros((64, 3))
        n = self._f["contact_forces"](self._h, i, f.ctypes.data, nrm.ctypes.data)
        return f[:n].copy(), nrm[:n].copy()

    def set_state(self, i, qpos, qvel):
        qpos = np.ascontiguousarray(qpos, dtype=np.float64); qvel = np.ascontiguousarray(qvel, dtype=np.float64)
        self._f["set_state"](self._h, i, qpos.ctypes.data, qvel.ctypes.data)

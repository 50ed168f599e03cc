"""Hopper-v5 and Walker2d-v5 on the B200 engine.

Mirror ``HopperEnv`` v5 (gymnasium/envs/mujoco/hopper_v5.py:24-343) and ``Walker2dEnv`` v5 (walker2d_v5.py:24-342) and the
``MujocoEnv`` plumbing they use (gymnasium/envs/mujoco/mujoco_env.py:35-229) behind the vector API with SyncVectorEnv's
conventions; the multibody dynamics the reference delegates to the MuJoCo wheel run in ``gymnasium_b200/csrc/mjc_planar.cuh``
(the articulated-body / PGS core of the Humanoid kernel generalised to slide + hinge joints).  Numeric parity with the real
MuJoCo wheel is unpinned (it cannot be installed here); see DESIGN.md.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .._api import AutoresetMode, Box
from ..vector_env import B200VectorEnv, ptr

INFO_KEYS = ("x_position", "z_distance_from_origin", "x_velocity", "reward_forward", "reward_ctrl", "reward_survive")


class _MjPlanarVectorEnv(B200VectorEnv):
    """Shared host logic of the planar MuJoCo robots: observation ``(N, 2 nq - 1) float64`` = qpos[1:] | clip(qvel, -10, 10),
    action ``(N, nu) float32`` in [-1, 1], reward float64, info = the keys of the env's ``step`` as ``(N,)`` arrays with
    their ``_key`` masks."""

    metadata = {"render_modes": [], "render_fps": 125, "autoreset_mode": AutoresetMode.NEXT_STEP}
    discrete_actions = False
    soa_output_keys = ("info",)
    robot, xml, NQ, NU = "hopper", "hopper.xml", 6, 3
    TIMESTEP, ACT_HIGH = 0.002, 1.0
    INFO_ROWS = {k: i for i, k in enumerate(INFO_KEYS)}  # key -> row of the kernel's info[6][n]
    RESET_KEYS, STEP_KEYS = INFO_KEYS[:2], INFO_KEYS[2:]  # _get_reset_info keys / keys only step() reports

    @classmethod
    def _obs_layout(cls):
        """(observation size, observation_structure): qpos[1:] | clip(qvel) for the locomotion robots (hopper_v5.py:241-246)."""
        return 2 * cls.NQ - 1, {"skipped_qpos": 1, "qpos": cls.NQ - 1, "qvel": cls.NQ}

    def __init__(self, num_envs, max_episode_steps, xml_file, frame_skip, forward_reward_weight, ctrl_cost_weight,
                 healthy_reward, terminate_when_unhealthy, healthy_state_range, healthy_z_range, healthy_angle_range,
                 reset_noise_scale, exclude_current_positions_from_observation, render_mode, engine_kwargs):
        if xml_file != self.xml:
            raise NotImplementedError(f"gymnasium_b200 compiles the stock {self.xml} only")
        if not exclude_current_positions_from_observation:
            raise NotImplementedError("only the default observation layout (x position excluded) is implemented")
        nq, nu = self.NQ, self.NU
        self.obs_size, structure = self._obs_layout()
        obs_space = Box(low=-np.inf, high=np.inf, shape=(self.obs_size,), dtype=np.float64)  # hopper_v5.py:237-239
        act_space = Box(low=-self.ACT_HIGH, high=self.ACT_HIGH, shape=(nu,), dtype=np.float32)  # ctrlrange, mujoco_env.py:105-110
        super().__init__(num_envs, obs_space, act_space, max_episode_steps=max_episode_steps, render_mode=render_mode,
                         **engine_kwargs)
        n, dev = self.num_envs, self.device
        self.frame_skip = int(frame_skip)
        self.dt = self.TIMESTEP * self.frame_skip  # mujoco_env.py:189-191
        self.observation_structure = structure
        self._cfg = _lib.MjPlanarCfg(
            reset_noise_scale=float(reset_noise_scale), forward_reward_weight=float(forward_reward_weight),
            ctrl_cost_weight=float(ctrl_cost_weight), healthy_reward=float(healthy_reward),
            healthy_z_min=float(healthy_z_range[0]), healthy_z_max=float(healthy_z_range[1]),
            healthy_angle_min=float(healthy_angle_range[0]), healthy_angle_max=float(healthy_angle_range[1]),
            healthy_state_min=float(healthy_state_range[0]), healthy_state_max=float(healthy_state_range[1]),
            terminate_when_unhealthy=int(bool(terminate_when_unhealthy)), frame_skip=self.frame_skip)
        self._s = {
            "qpos": torch.zeros((nq, n), dtype=torch.float64, device=dev),
            "qvel": torch.zeros((nq, n), dtype=torch.float64, device=dev),
            "qacc_warmstart": torch.zeros((nq, n), dtype=torch.float64, device=dev),
            "overflow": torch.zeros(1, dtype=torch.int32, device=dev),
        }
        self._last_done = None
        self._info_sid, self._info_prev = -1, None
        self._state = _lib.MjPlanarState(ctrl=self._ctrl.data_ptr(), rng=ptr(self._rng),
                                       **{k: v.data_ptr() for k, v in self._s.items()})

    def _alloc_outputs(self):
        n = self.num_envs
        layout = {"obs": ((n, self.obs_size), torch.float64), "reward": ((n,), torch.float64),
                  "info": ((len(INFO_KEYS), n), torch.float64), "terminated": ((n,), torch.bool),
                  "truncated": ((n,), torch.bool)}
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            layout["final_obs"] = ((n, self.obs_size), torch.float64)
        out = self._alloc_packed(layout)
        if "final_obs" in out:
            out["final_obs"].zero_()
        return out

    def _prepare_actions(self, actions):
        n = self.num_envs
        if isinstance(actions, torch.Tensor):
            t = actions
            if t.dim() != 2 or tuple(t.shape) != (n, self.NU):
                raise ValueError(f"Action dimension mismatch. Expected {(n, self.NU)}, found {tuple(t.shape)}")  # mujoco_env.py:198-201
            if t.dtype not in (torch.float32, torch.float64):
                t = t.to(torch.float32)
            return t.to(self.device).contiguous()
        a = np.asarray(actions)
        if a.ndim == 0:
            raise TypeError(f"actions must have a leading dimension of num_envs={n}, got a scalar")
        if a.shape != (n, self.NU):
            raise ValueError(f"Action dimension mismatch. Expected {(n, self.NU)}, found {a.shape}")
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float32)
        return super()._prepare_actions(np.ascontiguousarray(a))

    def _reset_kernel(self, mask, options, out):
        if mask is not None:
            if self.copy and self._has_reset:
                out["obs"].copy_(self._last_obs)
            out["info"].zero_()
        _lib.check(
            getattr(self._lib, f"b2e_{self.robot}_reset")(C.byref(self._batch), C.byref(self._cfg), C.byref(self._state),
                                       ptr(None if mask is None else mask.view(torch.uint8)), ptr(out["obs"]),
                                       ptr(out["info"]), self._stream),
            f"b2e_{self.robot}_reset",
        )
        self._last_obs = out["obs"]

    def _step_kernel(self, actions, out):
        _lib.check(
            getattr(self._lib, f"b2e_{self.robot}_step")(C.byref(self._batch), C.byref(self._cfg), C.byref(self._state), ptr(actions),
                                      ptr(out["obs"]), ptr(out["reward"]), ptr(out["terminated"]), ptr(out["truncated"]),
                                      ptr(out["info"]), ptr(out.get("final_obs")), self._stream),
            f"b2e_{self.robot}_step",
        )
        self._last_obs = out["obs"]

    def _info_dict(self, out, mask, keys):
        raw = out["info"]
        n = self.num_envs
        if mask is None:
            mask = np.ones(n, dtype=np.bool_) if isinstance(raw, np.ndarray) else torch.ones(n, dtype=torch.bool, device=self.device)
        return {k2: v for k in keys for k2, v in ((k, raw[self.INFO_ROWS[k]]), ("_" + k, mask))}

    def _same_kind(self, x, like):
        if isinstance(x, np.ndarray) == isinstance(like, np.ndarray):
            return x
        return x.cpu().numpy() if isinstance(x, torch.Tensor) else torch.as_tensor(x, device=self.device)

    def _reset_info(self, out, mask):
        # HopperEnv._get_reset_info (hopper_v5.py:339-343): the two position keys
        if mask is None:
            self._last_done = None
        elif self._last_done is not None:
            self._last_done = self._same_kind(self._last_done, mask) & ~mask
        self._info_sid = -1
        return self._info_dict(out, mask, self.RESET_KEYS)

    def _step_info(self, out):
        """Step info; lanes on their NEXT_STEP reset call (or, in SAME_STEP mode, lanes that just ended an episode) report
        only the reset-info keys, as ``SyncVectorEnv._add_info`` leaves them (see HumanoidVectorEnv._step_info)."""
        done = out["terminated"] | out["truncated"]
        sid = int(self._batch.call_counter)
        if sid != self._info_sid:
            self._info_sid, self._info_prev = sid, self._last_done
        info = self._info_dict(out, None, self.RESET_KEYS)
        step_mask = None
        if self.autoreset_mode == AutoresetMode.NEXT_STEP and self._info_prev is not None:
            step_mask = ~self._same_kind(self._info_prev, done)
        elif self.autoreset_mode == AutoresetMode.SAME_STEP:
            step_mask = ~done
        info.update(self._info_dict(out, step_mask, self.STEP_KEYS))
        self._last_done = done.copy() if isinstance(done, np.ndarray) else done
        if self.autoreset_mode == AutoresetMode.SAME_STEP:
            info.update({"final_obs": out["final_obs"], "_final_obs": done})
        return info

    # introspection -------------------------------------------------------------------------------------------------
    def qpos(self) -> torch.Tensor:
        return self._s["qpos"].t().contiguous()

    def qvel(self) -> torch.Tensor:
        return self._s["qvel"].t().contiguous()

    def buffer_overflow(self) -> bool:
        """True if any env ever exhausted the per-env contact (8) or constraint-row (16 Hopper / 32 Walker2d) buffers."""
        return bool(self._s["overflow"].item())


class HopperVectorEnv(_MjPlanarVectorEnv):
    """N Hopper-v5 envs (hopper_v5.py:163-178 for the keyword arguments)."""

    robot, xml, NQ, NU = "hopper", "hopper.xml", 6, 3
    TIMESTEP, ACT_HIGH = 0.002, 1.0
    INFO_ROWS = {k: i for i, k in enumerate(INFO_KEYS)}  # key -> row of the kernel's info[6][n]
    RESET_KEYS, STEP_KEYS = INFO_KEYS[:2], INFO_KEYS[2:]  # _get_reset_info keys / keys only step() reports

    @classmethod
    def _obs_layout(cls):
        """(observation size, observation_structure): qpos[1:] | clip(qvel) for the locomotion robots (hopper_v5.py:241-246)."""
        return 2 * cls.NQ - 1, {"skipped_qpos": 1, "qpos": cls.NQ - 1, "qvel": cls.NQ}

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 1000, xml_file: str = "hopper.xml",
                 frame_skip: int = 4, forward_reward_weight: float = 1.0, ctrl_cost_weight: float = 1e-3,
                 healthy_reward: float = 1.0, terminate_when_unhealthy: bool = True,
                 healthy_state_range=(-100.0, 100.0), healthy_z_range=(0.7, float("inf")),
                 healthy_angle_range=(-0.2, 0.2), reset_noise_scale: float = 5e-3,
                 exclude_current_positions_from_observation: bool = True, render_mode: str | None = None, **engine_kwargs):
        super().__init__(num_envs, max_episode_steps, xml_file, frame_skip, forward_reward_weight, ctrl_cost_weight,
                         healthy_reward, terminate_when_unhealthy, healthy_state_range, healthy_z_range, healthy_angle_range,
                         reset_noise_scale, exclude_current_positions_from_observation, render_mode, engine_kwargs)


class Walker2dVectorEnv(_MjPlanarVectorEnv):
    """N Walker2d-v5 envs (walker2d_v5.py:172-186 for the keyword arguments)."""

    robot, xml, NQ, NU = "walker2d", "walker2d_v5.xml", 9, 6

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 1000, xml_file: str = "walker2d_v5.xml",
                 frame_skip: int = 4, forward_reward_weight: float = 1.0, ctrl_cost_weight: float = 1e-3,
                 healthy_reward: float = 1.0, terminate_when_unhealthy: bool = True, healthy_z_range=(0.8, 2.0),
                 healthy_angle_range=(-1.0, 1.0), reset_noise_scale: float = 5e-3,
                 exclude_current_positions_from_observation: bool = True, render_mode: str | None = None, **engine_kwargs):
        super().__init__(num_envs, max_episode_steps, xml_file, frame_skip, forward_reward_weight, ctrl_cost_weight,
                         healthy_reward, terminate_when_unhealthy, (-np.inf, np.inf), healthy_z_range, healthy_angle_range,
                         reset_noise_scale, exclude_current_positions_from_observation, render_mode, engine_kwargs)


class InvertedPendulumVectorEnv(_MjPlanarVectorEnv):
    """N InvertedPendulum-v5 envs (inverted_pendulum_v5.py:110-116 for the keyword arguments): observation ``(N, 4) float64``
    = qpos | qvel, action ``(N, 1) float32`` in [-3, 3], reward 1.0 while the pole is within 0.2 rad (else 0.0 and terminated),
    info ``reward_survive``; reset info is empty."""

    robot, xml, NQ, NU = "inverted_pendulum", "inverted_pendulum.xml", 2, 1
    TIMESTEP, ACT_HIGH = 0.02, 3.0
    INFO_ROWS = {"reward_survive": 5}
    RESET_KEYS, STEP_KEYS = (), ("reward_survive",)
    metadata = {"render_modes": [], "render_fps": 25, "autoreset_mode": AutoresetMode.NEXT_STEP}

    @classmethod
    def _obs_layout(cls):
        return cls.NQ + cls.NQ, {"qpos": cls.NQ, "qvel": cls.NQ}  # inverted_pendulum_v5.py:140-143

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 1000, xml_file: str = "inverted_pendulum.xml",
                 frame_skip: int = 2, reset_noise_scale: float = 0.01, render_mode: str | None = None, **engine_kwargs):
        super().__init__(num_envs, max_episode_steps, xml_file, frame_skip, 0.0, 0.0, 0.0, True, (-np.inf, np.inf),
                         (-np.inf, np.inf), (-np.inf, np.inf), reset_noise_scale, True, render_mode, engine_kwargs)


class HalfCheetahVectorEnv(_MjPlanarVectorEnv):
    """N HalfCheetah-v5 envs (half_cheetah_v5.py:153-162 for the keyword arguments): observation ``(N, 17) float64`` =
    qpos[1:] | qvel (not clipped), action ``(N, 6) float32`` in [-1, 1], reward = forward_reward_weight * x_velocity -
    ctrl_cost_weight * sum(action^2); never terminates (TimeLimit 1000); reset noise: uniform on qpos, ``reset_noise_scale *
    standard_normal`` on qvel (numpy's ziggurat, restated on the device)."""

    robot, xml, NQ, NU = "half_cheetah", "half_cheetah.xml", 9, 6
    TIMESTEP, ACT_HIGH = 0.01, 1.0
    INFO_ROWS = {"x_position": 0, "x_velocity": 2, "reward_forward": 3, "reward_ctrl": 4}
    RESET_KEYS, STEP_KEYS = ("x_position",), ("x_velocity", "reward_forward", "reward_ctrl")
    metadata = {"render_modes": [], "render_fps": 20, "autoreset_mode": AutoresetMode.NEXT_STEP}

    def __init__(self, num_envs: int = 1, max_episode_steps: int | None = 1000, xml_file: str = "half_cheetah.xml",
                 frame_skip: int = 5, forward_reward_weight: float = 1.0, ctrl_cost_weight: float = 0.1,
                 reset_noise_scale: float = 0.1, exclude_current_positions_from_observation: bool = True,
                 render_mode: str | None = None, **engine_kwargs):
        super().__init__(num_envs, max_episode_steps, xml_file, frame_skip, forward_reward_weight, ctrl_cost_weight, 0.0, False,
                         (-np.inf, np.inf), (-np.inf, np.inf), (-np.inf, np.inf), reset_noise_scale,
                         exclude_current_positions_from_observation, render_mode, engine_kwargs)

"""Vector wrappers that stay on the device.

Torch restatements of the three stateful vector wrappers that sit directly on ``step()`` outputs in training loops:
``RecordEpisodeStatistics`` (gymnasium/wrappers/vector/common.py:22-235), ``NormalizeObservation``
(gymnasium/wrappers/vector/stateful_observation.py:22-164) and ``NormalizeReward``
(gymnasium/wrappers/vector/stateful_reward.py:21-178) with ``RunningMeanStd`` (gymnasium/wrappers/utils.py:33-71).
They consume and return torch tensors on the env's device and never synchronise with the host inside ``step()``
(the reference versions are NumPy and branch on ``np.sum(dones)`` every step).  Deviations, all for that reason:
``info["episode"]`` is present on every step (all-zero with an all-False ``_episode`` mask when nothing finished);
``episode_count`` and the queues live in device tensors and synchronise only when read.
"""
from __future__ import annotations

import time

import torch

from ._api import HAVE_GYMNASIUM, AutoresetMode

if HAVE_GYMNASIUM:
    from gymnasium.vector import VectorWrapper as _Base
else:  # pragma: no cover

    class _Base:
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        def reset(self, *, seed=None, options=None):
            return self.env.reset(seed=seed, options=options)

        def step(self, actions):
            return self.env.step(actions)

        def close(self, **kw):
            return self.env.close(**kw)


def _mode(env) -> AutoresetMode:
    m = env.metadata.get("autoreset_mode", AutoresetMode.NEXT_STEP)
    return m if isinstance(m, AutoresetMode) else AutoresetMode(getattr(m, "value", m))


def _device_of(env):
    return getattr(env.unwrapped if hasattr(env, "unwrapped") else env, "device", torch.device("cpu"))


class RunningMeanStd:
    """gymnasium/wrappers/utils.py:33-71 on torch tensors; ``count`` is a 0-dim tensor so masked updates need no sync."""

    def __init__(self, shape=(), dtype=torch.float64, device="cpu", epsilon=1e-4):
        self.mean = torch.zeros(shape, dtype=dtype, device=device)
        self.var = torch.ones(shape, dtype=dtype, device=device)
        self.count = torch.tensor(epsilon, dtype=torch.float64, device=device)

    def update(self, x: torch.Tensor, weight: torch.Tensor | None = None) -> None:
        """Batch update; ``weight`` (bool, per row) restricts it to a subset without a host round trip."""
        if weight is None:
            batch_mean = x.mean(dim=0)
            batch_var = x.var(dim=0, unbiased=False)
            batch_count = torch.tensor(float(x.shape[0]), dtype=torch.float64, device=x.device)
        else:
            w = weight.to(x.dtype)
            shape = (-1,) + (1,) * (x.dim() - 1)
            cnt = w.sum()
            safe = torch.clamp(cnt, min=1.0)
            batch_mean = (x * w.view(shape)).sum(dim=0) / safe
            batch_var = (((x - batch_mean) ** 2) * w.view(shape)).sum(dim=0) / safe
            batch_count = cnt.to(torch.float64)
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        dt = self.mean.dtype
        self.mean = (self.mean + delta * (batch_count / tot).to(dt)).to(dt)
        m_a = self.var * self.count.to(dt)
        m_b = batch_var * batch_count.to(dt)
        m2 = m_a + m_b + torch.square(delta) * (self.count * batch_count / tot).to(dt)
        self.var = (m2 / tot.to(dt)).to(dt)
        self.count = tot


class RecordEpisodeStatistics(_Base):
    """Cumulative reward / length / wall time per sub-env, reported in ``info[stats_key]`` when an episode ends."""

    def __init__(self, env, buffer_length: int = 100, stats_key: str = "episode"):
        super().__init__(env)
        self._stats_key = stats_key
        self._autoreset_mode = _mode(env)
        dev, n = _device_of(env), env.num_envs
        self._dev = dev
        self.episode_returns = torch.zeros(n, dtype=torch.float64, device=dev)
        self.episode_lengths = torch.zeros(n, dtype=torch.int64, device=dev)
        self.episode_start_times = torch.full((n,), time.perf_counter(), dtype=torch.float64, device=dev)
        self.prev_dones = torch.zeros(n, dtype=torch.bool, device=dev)
        self._count = torch.zeros((), dtype=torch.int64, device=dev)
        self._buffer_length = int(buffer_length)
        self._ring = {k: torch.zeros(self._buffer_length, dtype=dt, device=dev)
                      for k, dt in (("r", torch.float64), ("l", torch.int64), ("t", torch.float64))}

    @property
    def episode_count(self) -> int:
        return int(self._count.item())

    def _queue(self, key):
        """The reference's ``deque(maxlen=buffer_length)`` (wrappers/vector/common.py:103-105): oldest first."""
        count, B = self.episode_count, self._buffer_length
        if count <= B:
            return self._ring[key][:count]
        return torch.roll(self._ring[key], -(count % B))  # the ring's write position is the oldest entry

    @property
    def return_queue(self):
        return self._queue("r")

    @property
    def length_queue(self):
        return self._queue("l")

    @property
    def time_queue(self):
        return self._queue("t")

    def reset(self, *, seed=None, options=None):
        mask = None
        if options is not None and "reset_mask" in options:
            m = options["reset_mask"]
            mask = torch.as_tensor(m, dtype=torch.bool, device=self._dev) if not isinstance(m, torch.Tensor) else m.to(self._dev)
        obs, info = self.env.reset(seed=seed, options=options)
        now = time.perf_counter()
        if mask is None:
            mask = torch.ones_like(self.prev_dones)
        self.episode_start_times = torch.where(mask, torch.full_like(self.episode_start_times, now), self.episode_start_times)
        self.episode_returns = torch.where(mask, torch.zeros_like(self.episode_returns), self.episode_returns)
        self.episode_lengths = torch.where(mask, torch.zeros_like(self.episode_lengths), self.episode_lengths)
        self.prev_dones = self.prev_dones & ~mask
        return obs, info

    def step(self, actions):
        obs, rewards, term, trunc, infos = self.env.step(actions)
        rewards64 = rewards.to(torch.float64)
        now = time.perf_counter()
        if self._autoreset_mode == AutoresetMode.SAME_STEP:
            self.episode_returns = self.episode_returns + rewards64
            self.episode_lengths = self.episode_lengths + 1
        else:  # common.py:180-190: lanes whose previous call ended an episode are on their reset call
            live = ~self.prev_dones
            self.episode_returns = torch.where(live, self.episode_returns + rewards64, torch.zeros_like(rewards64))
            self.episode_lengths = torch.where(live, self.episode_lengths + 1, torch.zeros_like(self.episode_lengths))
            self.episode_start_times = torch.where(live, self.episode_start_times,
                                                   torch.full_like(self.episode_start_times, now))
        dones = term | trunc
        self.prev_dones = dones
        if self._stats_key in infos or f"_{self._stats_key}" in infos:
            raise ValueError(f"Attempted to add episode stats with key '{self._stats_key}' but this key already exists "
                             f"in info: {list(infos.keys())}")
        elapsed = torch.round((now - self.episode_start_times) * 1e6) / 1e6
        infos = dict(infos)
        infos[self._stats_key] = {
            "r": torch.where(dones, self.episode_returns, torch.zeros_like(self.episode_returns)),
            "l": torch.where(dones, self.episode_lengths, torch.zeros_like(self.episode_lengths)),
            "t": torch.where(dones, elapsed, torch.zeros_like(elapsed)),
        }
        infos[f"_{self._stats_key}"] = dones
        # ring buffers of finished episodes (the reference's deques), written at count + rank-among-dones
        rank = torch.cumsum(dones.to(torch.int64), 0) - 1
        n_done = dones.sum()
        keep = dones & (rank >= n_done - self._buffer_length)  # more finishers than slots: only the last ones survive (deque)
        slot = torch.where(keep, (self._count + rank) % self._buffer_length, torch.full_like(rank, self._buffer_length))
        for key, src in (("r", self.episode_returns), ("l", self.episode_lengths), ("t", elapsed)):
            padded = torch.cat([self._ring[key], self._ring[key].new_zeros(1)])
            padded[slot] = src.to(padded.dtype)
            self._ring[key] = padded[:-1]
        self._count = self._count + n_done
        if self._autoreset_mode == AutoresetMode.SAME_STEP:
            self.episode_returns = torch.where(dones, torch.zeros_like(self.episode_returns), self.episode_returns)
            self.episode_lengths = torch.where(dones, torch.zeros_like(self.episode_lengths), self.episode_lengths)
            self.episode_start_times = torch.where(dones, torch.full_like(self.episode_start_times, now),
                                                   self.episode_start_times)
        return obs, rewards, term, trunc, infos


class NormalizeObservation(_Base):
    """Running mean/variance normalisation of the batched observation (float32 statistics like the reference)."""

    def __init__(self, env, epsilon: float = 1e-8):
        if epsilon <= 0:
            raise ValueError(f"`epsilon` should be strictly positive. Received {epsilon}")
        super().__init__(env)
        if _mode(env) not in {AutoresetMode.NEXT_STEP}:
            raise ValueError(f"Expected env.metadata['autoreset_mode'] to be AutoresetMode.NEXT_STEP, got {_mode(env)}")
        shape = tuple(env.single_observation_space.shape)
        self.obs_rms = RunningMeanStd(shape=shape, dtype=torch.float32, device=_device_of(env))
        self.epsilon = float(epsilon)
        self.update_running_mean = True

    def observations(self, observations: torch.Tensor) -> torch.Tensor:
        x = observations.to(torch.float32)
        if self.update_running_mean:
            self.obs_rms.update(x)
        return ((x - self.obs_rms.mean) / torch.sqrt(self.obs_rms.var + self.epsilon)).to(torch.float32)

    def reset(self, *, seed=None, options=None):
        if options is not None and "reset_mask" in options:
            m = options["reset_mask"]
            if not bool(torch.as_tensor(m).all()):
                raise ValueError("NormalizeObservation does not support partial resets. The 'reset_mask' must contain all True values.")
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observations(obs), info

    def step(self, actions):
        obs, r, te, tr, info = self.env.step(actions)
        return self.observations(obs), r, te, tr, info


class NormalizeReward(_Base):
    """Scales rewards by the running standard deviation of the discounted return (stateful_reward.py:140-172)."""

    def __init__(self, env, gamma: float = 0.99, epsilon: float = 1e-8):
        if not 0 <= gamma <= 1:
            raise ValueError(f"`gamma` should be in the interval [0, 1]. Received {gamma}")
        if epsilon <= 0:
            raise ValueError(f"`epsilon` should be strictly positive. Received {epsilon}")
        super().__init__(env)
        dev = _device_of(env)
        self.return_rms = RunningMeanStd(shape=(), dtype=torch.float64, device=dev)
        self.accumulated_reward = torch.zeros(env.num_envs, dtype=torch.float32, device=dev)
        self.gamma, self.epsilon = float(gamma), float(epsilon)
        self.update_running_mean = True
        self._prev_dones = torch.zeros(env.num_envs, dtype=torch.bool, device=dev)
        self._autoreset_mode = _mode(env)

    def reset(self, *, seed=None, options=None):
        self.accumulated_reward.zero_()
        self._prev_dones.zero_()
        return self.env.reset(seed=seed, options=options)

    def step(self, actions):
        obs, reward, term, trunc, info = self.env.step(actions)
        active = torch.ones_like(self._prev_dones) if self._autoreset_mode == AutoresetMode.SAME_STEP else ~self._prev_dones
        r64 = reward.to(torch.float64)
        new = self.accumulated_reward.to(torch.float64) * self.gamma * (1 - term.to(torch.float64)) + r64
        self.accumulated_reward = torch.where(active, new.to(torch.float32), self.accumulated_reward)
        if self.update_running_mean:
            self.return_rms.update(self.accumulated_reward.to(torch.float64), weight=active)
        self._prev_dones = term | trunc
        if self._autoreset_mode == AutoresetMode.SAME_STEP:
            self.accumulated_reward = torch.where(self._prev_dones, torch.zeros_like(self.accumulated_reward),
                                                  self.accumulated_reward)
        return obs, r64 / torch.sqrt(self.return_rms.var + self.epsilon), term, trunc, info


class DictInfoToList(_Base):
    """``info`` dict of batched arrays + ``_key`` masks -> list of per-env dicts
    (gymnasium/wrappers/vector/dict_info_to_list.py:16-163).  The list is host data by definition, so tensors are
    brought to the host ONCE per call (one ``.cpu()`` per info array) instead of once per element; nested dicts and
    masks behave as in the reference.  Must be the outermost wrapper, like the reference's."""

    def reset(self, *, seed=None, options=None):
        obs, infos = self.env.reset(seed=seed, options=options)
        return obs, self._convert(infos)

    def step(self, actions):
        obs, reward, terminated, truncated, infos = self.env.step(actions)
        return obs, reward, terminated, truncated, self._convert(infos)

    def _convert(self, vector_infos):
        assert isinstance(vector_infos, dict)
        n = self.num_envs
        out = [{} for _ in range(n)]
        for key, value in vector_infos.items():
            if key.startswith("_"):
                continue
            mask = vector_infos.get(f"_{key}")
            if isinstance(mask, torch.Tensor):
                mask = mask.cpu().numpy()
            if isinstance(value, dict):
                value = self._convert(value)
            elif isinstance(value, torch.Tensor):
                value = value.cpu().numpy()
            assert len(value) == n, f"Expects {key!r} to have length equal to the num-envs ({n}), actual length is {len(value)}"
            if mask is not None:
                assert len(mask) == n, f"Expects {'_' + key!r} to have length equal to the num-envs ({n}), actual length is {len(mask)}"
            for i in range(n):
                if mask is None or mask[i]:
                    out[i][key] = value[i]
        return out


class NumpyToTorch(_Base):
    """Drop-in for ``gymnasium.wrappers.vector.NumpyToTorch`` (numpy_to_torch.py:15-80) around an engine env.

    The reference wrapper converts a NumPy env's outputs to torch tensors (and torch actions back to NumPy) on every call.
    An engine env already computes on the device: with ``output="torch"`` this wrapper is a no-op pass-through (actions
    may be torch tensors on any device or numpy arrays; nothing is copied to the host), and with ``output="numpy"`` it
    switches the env to torch outputs instead of converting back and forth.  ``device`` moves the outputs if it differs
    from the env's device."""

    def __init__(self, env, device=None):
        super().__init__(env)
        base = getattr(env, "unwrapped", env)
        if getattr(base, "output", "torch") == "numpy":
            base.output = "torch"  # stop producing host arrays that would only be converted back
        self.device = None if device is None else torch.device(device)

    def _out(self, x):
        if isinstance(x, dict):
            return {k: self._out(v) for k, v in x.items()}
        if isinstance(x, tuple):
            return tuple(self._out(v) for v in x)
        if not isinstance(x, torch.Tensor):
            try:
                x = torch.as_tensor(x)
            except (TypeError, ValueError, RuntimeError):
                return x
        return x if self.device is None or x.device == self.device else x.to(self.device)

    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self._out(obs), self._out(info)

    def step(self, actions):
        obs, reward, terminated, truncated, info = self.env.step(actions)
        return self._out(obs), self._out(reward), self._out(terminated), self._out(truncated), self._out(info)

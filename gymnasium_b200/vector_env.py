"""``B200VectorEnv`` -- the VectorEnv the engine exposes to Gymnasium.

Host-side mirror of ``gymnasium.vector.SyncVectorEnv`` (gymnasium/vector/sync_vector_env.py:76-403) for the
step()/reset() path: same constructor-visible attributes (``num_envs``, batched and single spaces,
``metadata["autoreset_mode"]``), same ``reset(seed=..., options=...)`` / ``step(actions)`` signatures, dtypes, seeding
convention (sub-env i gets ``seed + i``, :205-208) and error behaviour (:196-231), but every sub-env lives in a
struct-of-arrays batch in HBM and one CUDA launch (libb200env.so, include/b200env.h) advances all of them.

Returned arrays are torch tensors on the env's CUDA device (``output="torch"``, default) or host numpy arrays
(``output="numpy"``, one device->host copy per call).  There is no CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import secrets
from typing import Any

import numpy as np
import torch

from . import _lib
from ._api import AutoresetMode, VectorEnv, batch_space

_MODE = {
    AutoresetMode.NEXT_STEP: _lib.AUTORESET_NEXT_STEP,
    AutoresetMode.SAME_STEP: _lib.AUTORESET_SAME_STEP,
    AutoresetMode.DISABLED: _lib.AUTORESET_DISABLED,
}
_ACT_DTYPES = {torch.int64: _lib.ACT_I64, torch.int32: _lib.ACT_I32, torch.uint8: _lib.ACT_U8,
               torch.float32: _lib.ACT_F32, torch.float64: _lib.ACT_F64}
_U64 = (1 << 64) - 1


def _as_autoreset_mode(mode) -> AutoresetMode:
    if isinstance(mode, AutoresetMode):
        return mode
    for m in AutoresetMode:  # accept the enum of the *other* namespace (compat vs gymnasium) and plain strings
        if getattr(mode, "value", mode) == m.value:
            return m
    raise ValueError(f"Unexpected autoreset mode, {mode}")


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


class B200VectorEnv(VectorEnv):
    """Family-independent host logic; subclasses provide the spaces, device state and the two kernel calls."""

    metadata: dict[str, Any] = {"render_modes": [], "autoreset_mode": AutoresetMode.NEXT_STEP}
    discrete_actions = True  # Box-action families override

    def __init__(
        self,
        num_envs: int,
        single_observation_space,
        single_action_space,
        *,
        max_episode_steps: int | None,
        autoreset_mode=AutoresetMode.NEXT_STEP,
        device: str | int | torch.device | None = None,
        rng: str = "numpy",
        env_offset: int = 0,
        output: str = "torch",
        copy: bool = True,
        validate_actions: bool = False,
        out_buffers: int = 1,
        render_mode: str | None = None,
    ):
        if render_mode is not None:
            raise ValueError("gymnasium_b200 environments do not render (render_mode must be None)")
        if int(num_envs) <= 0:
            raise ValueError(f"num_envs must be positive, got {num_envs}")
        if rng not in ("numpy", "philox"):
            raise ValueError(f"rng must be 'numpy' (bit-exact PCG64 streams) or 'philox' (stateless), got {rng!r}")
        if output not in ("torch", "numpy"):
            raise ValueError(f"output must be 'torch' or 'numpy', got {output!r}")
        self.num_envs = int(num_envs)
        self.single_observation_space = single_observation_space
        self.single_action_space = single_action_space
        self.observation_space = batch_space(single_observation_space, self.num_envs)
        self.action_space = batch_space(single_action_space, self.num_envs)
        self.autoreset_mode = _as_autoreset_mode(autoreset_mode)
        self.metadata = {**type(self).metadata, "autoreset_mode": self.autoreset_mode}
        self.max_episode_steps = None if max_episode_steps is None else int(max_episode_steps)
        self.render_mode = None
        self.rng_mode = rng
        self.env_offset = int(env_offset)
        self.output = output
        self.copy = bool(copy)
        # copy=False only: rotate over this many preallocated output sets, so that a consumer (a D2H copy or a gather on
        # another stream) may still be reading step k's outputs while the kernel of step k+1 writes the next set
        self.out_buffers = max(1, int(out_buffers))
        self._out_ring: list[dict[str, torch.Tensor]] = []
        self._out_pos = 0
        # opt-in: the reference asserts `action_space.contains(action)` per sub-env (e.g. cartpole.py:165-167); checking
        # on the device costs a reduction + a host sync per step, so by default the kernels clamp instead
        self.validate_actions = bool(validate_actions)

        # ---- device + library: no fallback -------------------------------------------------------------------
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError(
                "gymnasium_b200 needs a CUDA device (B200, sm_100a): torch.cuda.is_available() is False and the "
                "engine has no CPU fallback. Use gymnasium.vector.SyncVectorEnv on CPU-only hosts."
            )
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise ValueError(f"device must be a CUDA device, got {dev}")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev

        n = self.num_envs
        self._batch = _lib.Batch(
            n=n,
            env_offset=self.env_offset,
            max_episode_steps=self.max_episode_steps if self.max_episode_steps is not None else 0,
            autoreset_mode=_MODE[self.autoreset_mode],
            rng_mode=_lib.RNG_NUMPY if rng == "numpy" else _lib.RNG_PHILOX,
            action_dtype=_lib.ACT_I64,
            philox_seed=0,
            call_counter=0,
        )
        self._ctrl = torch.zeros(n, dtype=torch.int32, device=dev)
        # numpy-parity streams: uint64 [2][n][2] = PCG64 state then increment (stored as int64 bit patterns)
        self._rng = torch.zeros((2, n, 2), dtype=torch.int64, device=dev) if rng == "numpy" else None
        self._seeded = False
        self._base_seed: int | None = None
        self._seed_list: list[int] | None = None
        self._has_reset = False
        self._pending: tuple | None = None  # ("reset" | "step", result, reset mask) between *_async and *_wait
        self._pinned_actions: list | None = None  # ring of (pinned tensor, numpy view, event of the last H2D out of it)
        self._pinned_pos = 0
        self._out: dict[str, torch.Tensor] = {}
        self._pinned_wire: torch.Tensor | None = None
        self._pinned_np: np.ndarray | None = None

    # ------------------------------------------------------------------------------------------------------------
    # hooks for families
    def _alloc_outputs(self) -> dict[str, torch.Tensor]:
        raise NotImplementedError

    def _reset_kernel(self, mask: torch.Tensor | None, options: dict | None, out: dict[str, torch.Tensor]) -> None:
        raise NotImplementedError

    def _step_kernel(self, actions: torch.Tensor, out: dict[str, torch.Tensor]) -> None:
        raise NotImplementedError

    def _reset_info(self, out, mask) -> dict:
        return {}

    def _step_info(self, out) -> dict:
        return {}

    # ------------------------------------------------------------------------------------------------------------
    @property
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _outputs(self) -> dict[str, torch.Tensor]:
        if self.copy:
            self._out = self._alloc_outputs()  # fresh tensors from the caching allocator; kernels write in place
        elif self.out_buffers > 1:
            if len(self._out_ring) < self.out_buffers:
                self._out_ring.append(self._alloc_outputs())
                self._out = self._out_ring[-1]
            else:
                self._out = self._out_ring[self._out_pos % self.out_buffers]
            self._out_pos += 1
        elif not self._out:
            self._out = self._alloc_outputs()
        return self._out

    def _alloc_packed(self, layout: dict[str, tuple[tuple, torch.dtype]]) -> dict[str, torch.Tensor]:
        """Allocates the per-call outputs as typed views of ONE contiguous byte buffer (widest dtype first, so every
        view is naturally aligned).  Kernels write through the views; ``output="numpy"`` then needs a single
        device->host copy for the whole step result."""
        items = sorted(layout.items(), key=lambda kv: -torch.empty((), dtype=kv[1][1]).element_size())
        sizes = [(k, shape, dt, int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()) for k, (shape, dt) in items]
        offsets, total = {}, 0
        for k, shape, dt, nbytes in sizes:
            total = (total + 15) // 16 * 16
            offsets[k] = total
            total += nbytes
        wire = torch.empty(total, dtype=torch.uint8, device=self.device)
        out = {"_wire": wire}
        for k, shape, dt, nbytes in sizes:
            raw = wire[offsets[k]:offsets[k] + nbytes]
            out[k] = (raw.view(torch.bool) if dt == torch.bool else raw.view(dt)).view(shape)
        self._wire_layout = [(k, shape, dt, offsets[k], nbytes) for k, shape, dt, nbytes in sizes]
        return out

    def packed_outputs(self) -> tuple[torch.Tensor, list]:
        """The last call's outputs as ONE contiguous uint8 device buffer + its layout
        ``[(key, shape, dtype, byte_offset, nbytes), ...]`` -- what a multi-GPU gather should move (one collective)."""
        return self._out["_wire"], list(self._wire_layout)

    def _to_host(self, out: dict[str, torch.Tensor]) -> dict[str, np.ndarray]:
        """One async D2H copy of the packed outputs into a cached pinned buffer + one stream sync -> numpy views."""
        wire = out["_wire"]
        if self._pinned_wire is None or self._pinned_wire.numel() != wire.numel():
            self._pinned_wire = torch.empty(wire.numel(), dtype=torch.uint8, pin_memory=True)
            self._pinned_np = self._pinned_wire.numpy()
        self._pinned_wire.copy_(wire, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        host = {}
        for k, shape, dt, off, nbytes in self._wire_layout:
            npdt = np.bool_ if dt == torch.bool else torch.empty((), dtype=dt).numpy().dtype
            a = self._pinned_np[off:off + nbytes].view(npdt).reshape(shape)
            host[k] = a.copy() if self.copy else a
        return host

    # ------------------------------------------------------------------------------------------------------------
    def _seed_streams(self, seed, mask: torch.Tensor | None) -> None:
        """SyncVectorEnv.reset seeding (sync_vector_env.py:203-212): int -> seed+i; list -> per env; None -> keep."""
        n = self.num_envs
        seeds_dev = None
        base = 0
        if seed is None:
            if self._seeded:
                return  # Env.reset(seed=None) keeps the stream (gymnasium/core.py:157-159)
            base = secrets.randbits(62)  # lazy self-seeding (core.py:226-235); kept for np_random_seed
            self._base_seed, self._seed_list = base, None
        elif isinstance(seed, (int, np.integer)):
            base = int(seed)
            if base < 0:
                raise ValueError(f"Seed must be a non-negative integer, got {base}")  # seeding.py:28-34
            if base + self.env_offset + n - 1 > _U64:
                raise ValueError("gymnasium_b200 supports seeds below 2**64")
            if mask is None or not self._seeded:
                self._base_seed, self._seed_list = base, None
            else:  # masked re-seed: only those lanes change their seed (np_random_seed stays per-lane exact)
                seeds = list(self.np_random_seed)
                for i in torch.nonzero(mask).flatten().tolist():
                    seeds[i] = base + self.env_offset + i
                self._seed_list, self._base_seed = seeds, None
        else:
            seed = list(seed)
            if len(seed) != n:
                raise ValueError(
                    f"If seeds are passed as a list the length must match num_envs={n} but got length={len(seed)}."
                )
            if any(s is None for s in seed):
                # per-env None: those lanes keep their stream (or get fresh entropy if never seeded)
                keep = torch.tensor([s is not None for s in seed], dtype=torch.bool)
                fill = [secrets.randbits(62) if (s is None and not self._seeded) else (0 if s is None else int(s))
                        for s in seed]
                lane_mask = keep if self._seeded else torch.ones(n, dtype=torch.bool)
                lane_mask = lane_mask.to(self.device)
                mask = lane_mask if mask is None else (mask & lane_mask)
                seed = fill
            if any(int(s) < 0 or int(s) > _U64 for s in seed):
                raise ValueError("gymnasium_b200 supports seeds in [0, 2**64)")
            arr = np.array([int(s) for s in seed], dtype=np.uint64).view(np.int64)
            seeds_dev = torch.from_numpy(arr).to(self.device)
            if mask is not None and self._seeded:  # only the masked lanes take their new seed
                old, m = list(self.np_random_seed), mask.cpu().tolist()
                self._seed_list, self._base_seed = [int(s) if m[i] else old[i] for i, s in enumerate(seed)], None
            else:
                self._seed_list, self._base_seed = [int(s) for s in seed], None
        if self.rng_mode == "numpy":
            if not self._seeded:
                # first seeding of this batch: EVERY lane gets its stream, whatever the reset mask says -- an unmasked lane
                # is an env whose np_random self-seeds lazily at its first draw (gymnasium/core.py:226-235); all-zero PCG64
                # words would make every later autoreset of that lane draw u = 0.0
                if mask is not None and seeds_dev is None:
                    mask = None
                elif mask is not None:
                    lazy = torch.from_numpy(np.array([secrets.randbits(62) for _ in range(n)], dtype=np.int64)).to(self.device)
                    seeds_dev = torch.where(mask, seeds_dev, lazy)
                    mask = None
            mask_u8 = None if mask is None else mask.view(torch.uint8)
            _lib.check(
                self._lib.b2e_rng_seed(C.byref(self._batch), base & _U64, ptr(seeds_dev), ptr(mask_u8),
                                       ptr(self._rng), self._stream),
                "b2e_rng_seed",
            )
            self._on_streams_seeded(mask)
        else:
            self._batch.philox_seed = (base if seeds_dev is None else int(seed[0])) & _U64
        self._seeded = True

    def _on_streams_seeded(self, lanes: torch.Tensor | None) -> None:
        """Hook: the numpy-parity streams of `lanes` (bool mask, None = all) were just (re)seeded; families that keep extra
        generator state next to the PCG64 words (Blackjack: numpy's 32-bit word buffer) reset it here."""

    def _parse_reset_mask(self, options: dict | None):
        """Validation and messages of sync_vector_env.py:214-231 (torch bool tensors are accepted too)."""
        if options is None or "reset_mask" not in options:
            return options, None
        options = dict(options)
        reset_mask = options.pop("reset_mask")
        n = self.num_envs
        if isinstance(reset_mask, torch.Tensor):
            if tuple(reset_mask.shape) != (n,):
                raise ValueError(f"`options['reset_mask']` must have shape `({n},)`, got {tuple(reset_mask.shape)}")
            if reset_mask.dtype != torch.bool:
                raise TypeError(f"`options['reset_mask']` must have `dtype=np.bool_`, got {reset_mask.dtype}")
            mask = reset_mask.to(self.device)
            if not bool(mask.any()):
                raise ValueError(
                    f"`options['reset_mask']` must contain a boolean array with at least one True value, got reset_mask={reset_mask}"
                )
            return options, mask.contiguous()
        if not isinstance(reset_mask, np.ndarray):
            raise TypeError(f"`options['reset_mask']` must be a numpy array, got {type(reset_mask)}")
        if reset_mask.shape != (n,):
            raise ValueError(f"`options['reset_mask']` must have shape `({n},)`, got {reset_mask.shape}")
        if reset_mask.dtype != np.bool_:
            raise TypeError(f"`options['reset_mask']` must have `dtype=np.bool_`, got {reset_mask.dtype}")
        if not np.any(reset_mask):
            raise ValueError(
                f"`options['reset_mask']` must contain a boolean array with at least one True value, got reset_mask={reset_mask}"
            )
        return options, torch.from_numpy(np.ascontiguousarray(reset_mask)).to(self.device)

    def reset(self, *, seed: int | list[int | None] | None = None, options: dict[str, Any] | None = None):
        """SyncVectorEnv.reset (gymnasium/vector/sync_vector_env.py:187-264) on device."""
        self._assert_open()
        options, mask = self._parse_reset_mask(options)
        with torch.cuda.device(self.device):
            self._seed_streams(seed, mask)
            out = self._outputs()
            self._reset_kernel(mask, options, out)
            self._batch.call_counter += 1
        self._has_reset = True
        if self.output == "numpy":
            host = self._to_host(out)
            m = None if mask is None else mask.cpu().numpy()
            return host["obs"], self._reset_info(host, m)
        return out["obs"], self._reset_info(out, mask)

    def _prepare_actions(self, actions) -> torch.Tensor:
        """One conversion to a contiguous device tensor; wrong count -> ValueError, scalar -> TypeError
        (tests/vector/test_vector_env.py:330-364)."""
        n = self.num_envs
        if isinstance(actions, torch.Tensor) and actions.device.type == "cpu":
            actions = actions.numpy()
        if isinstance(actions, torch.Tensor):
            t = actions
        else:
            if np.ndim(actions) == 0:
                raise TypeError(f"actions must be an iterable of length num_envs={n}, got a scalar {actions!r}")
            a = np.ascontiguousarray(actions)
            if a.dtype == object or a.dtype.kind not in "iufb":
                raise TypeError(f"unsupported action array dtype {a.dtype}")
            if a.dtype == np.bool_:
                a = a.astype(np.uint8)
            if self.discrete_actions and a.dtype.kind == "f":
                raise TypeError(f"discrete actions must be integers, got dtype {a.dtype}")
            if self.discrete_actions and a.dtype not in (np.int64, np.int32, np.uint8):
                a = a.astype(np.int64)
            if a.shape[0] != n:
                raise ValueError(f"expected {n} actions (one per sub-environment), got {a.shape[0]}")
            if self.discrete_actions and a.ndim != 1:
                raise ValueError(f"discrete actions must have shape ({n},), got {a.shape}")
            # stage through a small ring of cached pinned buffers with a plain single-threaded memcpy, then one async
            # H2D copy (a ring, so that the host may run ahead of the stream by a few steps without waiting for a DMA)
            ring = self._pinned_actions
            if ring is None or tuple(ring[0][0].shape) != a.shape or ring[0][1].dtype != a.dtype:
                ring = self._pinned_actions = []
                for _ in range(4):
                    pt = torch.from_numpy(np.empty(a.shape, dtype=a.dtype)).pin_memory()
                    ring.append([pt, pt.numpy(), None])
            slot = ring[self._pinned_pos % len(ring)]
            self._pinned_pos += 1
            if slot[2] is not None:
                slot[2].synchronize()  # the DMA that last read this staging buffer must have finished
            np.copyto(slot[1], a)
            t = slot[0].to(self.device, non_blocking=True)
            if slot[2] is None:
                slot[2] = torch.cuda.Event()
            slot[2].record(torch.cuda.current_stream(self.device))
            return t
        if t.dim() == 0:
            raise TypeError(f"actions must have a leading dimension of num_envs={n}, got a scalar tensor")
        if t.shape[0] != n:
            raise ValueError(f"expected {n} actions (one per sub-environment), got {t.shape[0]}")
        if self.discrete_actions:
            if t.dim() != 1:
                raise ValueError(f"discrete actions must have shape ({n},), got {tuple(t.shape)}")
            if t.dtype not in (torch.int64, torch.int32, torch.uint8):
                t = t.to(torch.int64)
        if t.device != self.device:
            t = t.to(self.device)
        return t.contiguous()

    def _check_actions(self, t: torch.Tensor) -> None:
        space = self.single_action_space
        if self.discrete_actions:
            bad = (t < int(getattr(space, "start", 0))) | (t >= int(getattr(space, "start", 0)) + int(space.n))
        else:
            lo = torch.as_tensor(space.low, device=t.device, dtype=t.dtype)
            hi = torch.as_tensor(space.high, device=t.device, dtype=t.dtype)
            bad = ((t < lo) | (t > hi) | torch.isnan(t)).reshape(t.shape[0], -1).any(dim=1)
        if bool(bad.any()):
            i = int(torch.nonzero(bad)[0])
            raise AssertionError(f"{t[i].tolist()!r} ({t.dtype}) invalid (sub-environment {i})")

    def step(self, actions):
        """SyncVectorEnv.step (gymnasium/vector/sync_vector_env.py:266-337) as one fused launch."""
        self._assert_open()
        if not self._has_reset:
            # OrderEnforcing.step (gymnasium/wrappers/common.py:393-397) raises ResetNeeded per sub-env
            from . import errors

            raise errors.ResetNeeded("Cannot call env.step() before calling env.reset()")
        with torch.cuda.device(self.device):
            t = self._prepare_actions(actions)
            if self.validate_actions:
                self._check_actions(t)
            out = self._outputs()
            self._batch.action_dtype = _ACT_DTYPES[t.dtype]
            if self.rng_mode == "philox" and torch.cuda.is_current_stream_capturing():
                # the call counter that keys this call's Philox draws is a kernel ARGUMENT: a graph replay would re-use it
                raise RuntimeError("rng='philox' cannot be captured in a CUDA graph (every replay would repeat the captured "
                                   "call counter and with it the reset draws); use rng='numpy' (per-env PCG64 state in HBM)")
            self._step_kernel(t, out)
            self._batch.call_counter += 1
        if self.output == "numpy":
            out = self._to_host(out)
        return out["obs"], out["reward"], out["terminated"], out["truncated"], self._step_info(out)

    # ------------------------------------------------------------------------------------------------------------
    @property
    def np_random_seed(self) -> tuple[int, ...]:
        """Seeds of the sub-envs, as SyncVectorEnv.np_random_seed (sync_vector_env.py:177-180)."""
        if not self._seeded:
            with torch.cuda.device(self.device):
                self._seed_streams(None, None)
        if self._seed_list is not None:
            return tuple(self._seed_list)
        return tuple(self._base_seed + self.env_offset + i for i in range(self.num_envs))

    @property
    def np_random(self) -> tuple:
        """``SyncVectorEnv.np_random`` (sync_vector_env.py:182-185): one ``numpy.random.Generator`` per sub-env.  The live
        streams are the PCG64 words on the device; these Generators are SNAPSHOTS of them (same state and increment, so they
        produce exactly the draws the env would make next) -- drawing from them does not advance the env's streams."""
        words = self.rng_state()
        if words is None:
            raise AttributeError("rng='philox' keeps no per-env generator state (stateless counter-based draws)")
        if not self._seeded:
            with torch.cuda.device(self.device):
                self._seed_streams(None, None)
            words = self.rng_state()
        gens = []
        for i in range(self.num_envs):
            bg = np.random.PCG64()
            bg.state = {"bit_generator": "PCG64",
                        "state": {"state": int(words[0, i, 0]) | (int(words[0, i, 1]) << 64),
                                  "inc": int(words[1, i, 0]) | (int(words[1, i, 1]) << 64)},
                        "has_uint32": 0, "uinteger": 0}
            gens.append(np.random.Generator(bg))
        return tuple(gens)

    def rng_state(self) -> np.ndarray | None:
        """uint64 [2][n][2] host copy of the PCG64 (state, inc) words; None in philox mode."""
        if self._rng is None:
            return None
        return self._rng.cpu().numpy().view(np.uint64)

    def elapsed_steps(self) -> torch.Tensor:
        """TimeLimit counters (gymnasium/wrappers/common.py:108) as an int32 device tensor."""
        return self._ctrl & 0x7FFFFFFF

    def close_extras(self, **kwargs):
        self._out = {}
        self._out_ring = []
        self._rng = None
        self._ctrl = None

    # SyncVectorEnv.call/get_attr/set_attr (sync_vector_env.py:343-398): there are no sub-env objects here
    def get_attr(self, name: str):
        if hasattr(self, name):
            v = getattr(self, name)
            return tuple(v for _ in range(self.num_envs)) if not isinstance(v, (torch.Tensor, tuple)) else v
        raise AttributeError(f"{type(self).__name__} has no per-env attribute {name!r}")

    def call(self, name: str, *args, **kwargs) -> tuple:
        """``SyncVectorEnv.call`` (sync_vector_env.py:343-367): one result per sub-env.  A method of the batched env is called
        once and its result repeated; per-env tensors / tuples are returned as they are."""
        attr = getattr(self, name, None)
        if attr is None:
            raise AttributeError(f"{type(self).__name__} has no attribute {name!r}")
        if callable(attr):
            v = attr(*args, **kwargs)
            return v if isinstance(v, (tuple, torch.Tensor)) and len(v) == self.num_envs else tuple(v for _ in range(self.num_envs))
        return self.get_attr(name)

    def set_attr(self, name: str, values) -> None:
        """``SyncVectorEnv.set_attr`` (sync_vector_env.py:380-398) for the scalar parameters of a family: every sub-env shares
        one kernel configuration here, so all values must agree (per-env parameters would need per-env kernel arguments)."""
        if not isinstance(values, (list, tuple)):
            values = [values for _ in range(self.num_envs)]
        if len(values) != self.num_envs:
            raise ValueError(
                "Values must be a list or tuple with length equal to the number of environments. "
                f"Got `{len(values)}` values for {self.num_envs} environments."
            )
        if any(v != values[0] for v in values[1:]):
            raise ValueError(f"{type(self).__name__}.set_attr needs the same value for every sub-environment (one kernel "
                             f"configuration per batch), got {values!r}")
        if not hasattr(self, name):
            raise AttributeError(f"{type(self).__name__} has no attribute {name!r}")
        setattr(self, name, values[0])
        cfg = getattr(self, "_cfg", None)
        if cfg is not None and any(f[0] == name for f in type(cfg)._fields_):
            setattr(cfg, name, values[0])

    # ---- AsyncVectorEnv's split calls (async_vector_env.py:310-521).  A kernel launch IS asynchronous: *_async enqueues the
    # work on the stream and returns, *_wait hands out the result (and, for output="numpy", is where the host copy is made)
    def _assert_open(self):
        if getattr(self, "closed", False):
            from . import errors

            raise errors.ClosedEnvironmentError(f"Trying to operate on `{type(self).__name__}`, after a call to `close()`.")

    def reset_async(self, seed=None, options=None) -> None:
        from . import errors

        self._assert_open()
        if self._pending is not None:
            raise errors.AlreadyPendingCallError(
                f"Calling `reset_async` while waiting for a pending call to `{self._pending[0]}` to complete", self._pending[0])
        output, self.output = self.output, "torch"
        try:
            res = self.reset(seed=seed, options=options)
        finally:
            self.output = output
        mask = None if options is None else options.get("reset_mask")
        self._pending = ("reset", res, mask)

    def reset_wait(self, timeout: float | None = None):
        from . import errors

        self._assert_open()
        if self._pending is None or self._pending[0] != "reset":
            raise errors.NoAsyncCallError("Calling `reset_wait` without any prior call to `reset_async`.", "reset")
        _, (obs, info), mask = self._pending
        self._pending = None
        if self.output == "numpy":
            host = self._to_host(self._out)  # the buffers the pending launch wrote
            m = None if mask is None else np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask)
            return self._host_obs(host), self._reset_info(host, m)
        return obs, info

    def step_async(self, actions) -> None:
        from . import errors

        self._assert_open()
        if self._pending is not None:
            raise errors.AlreadyPendingCallError(
                f"Calling `step_async` while waiting for a pending call to `{self._pending[0]}` to complete.", self._pending[0])
        output, self.output = self.output, "torch"
        try:
            res = self.step(actions)
        finally:
            self.output = output
        self._pending = ("step", res, None)

    def step_wait(self, timeout: float | None = None):
        from . import errors

        self._assert_open()
        if self._pending is None or self._pending[0] != "step":
            raise errors.NoAsyncCallError("Calling `step_wait` without any prior call to `step_async`.", "step")
        _, res, _ = self._pending
        self._pending = None
        if self.output == "numpy":
            host = self._to_host(self._out)
            return self._host_obs(host), host["reward"], host["terminated"], host["truncated"], self._step_info(host)
        return res

    def _host_obs(self, host):
        """Observation as the family's reset()/step() return it (overridden where that is not the raw array)."""
        return host["obs"]

"""Sharding a vector env across the GPUs of one box (one process per GPU, ``torch.distributed``).

Sub-envs are independent (no cross-env term anywhere on the path), so the global index range is cut into contiguous
shards, each rank steps its own shard with no exchange, and the only collective is an optional gather of the step
outputs into a single batch (NCCL over NVLink on GPUs; gloo on CPU in the tests).  Seeds use the GLOBAL env index
(``seed + env_offset + i``) so results do not depend on the number of ranks.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_bounds(total: int, world_size: int, rank: int) -> tuple[int, int]:
    """(start, count) of rank's contiguous shard; the first ``total % world_size`` ranks get one extra env."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(int(total), int(world_size))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults otherwise)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_sharded(env_id: str, total_envs: int, rank: int | None = None, world_size: int | None = None, **kwargs):
    """This rank's shard of a ``total_envs``-wide vector env (device = ``cuda:LOCAL_RANK`` unless given)."""
    from .registration import make_vec

    r, lr, w = env_rank_world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    start, count = shard_bounds(total_envs, world_size, rank)
    kwargs.setdefault("device", f"cuda:{lr}")
    return make_vec(env_id, num_envs=count, env_offset=start, **kwargs)


class BatchGather:
    """Gathers per-rank step outputs into one ``[total, ...]`` batch on every rank (``dst=None``) or on one rank.

    Equal shard sizes use ``all_gather_into_tensor`` (one NCCL call per tensor, written in place into the cached
    global batch); ragged shards are padded to the largest shard, gathered and compacted.
    """

    def __init__(self, total: int, world_size: int, rank: int, group=None, dst: int | None = None):
        self.total, self.world, self.rank, self.group, self.dst = int(total), world_size, rank, group, dst
        self.bounds = [shard_bounds(total, world_size, r) for r in range(world_size)]
        self.equal = len({c for _, c in self.bounds}) == 1
        self._buf: dict[str, torch.Tensor] = {}

    def _buffer(self, key: str, like: torch.Tensor) -> torch.Tensor:
        shape = (self.total,) + tuple(like.shape[1:])
        b = self._buf.get(key)
        if b is None or b.shape != shape or b.dtype != like.dtype or b.device != like.device:
            b = torch.empty(shape, dtype=like.dtype, device=like.device)
            self._buf[key] = b
        return b

    def __call__(self, **tensors: torch.Tensor) -> dict[str, torch.Tensor]:
        out = {}
        for key, t in tensors.items():
            if t.shape[0] != self.bounds[self.rank][1]:
                raise ValueError(f"{key}: leading dim {t.shape[0]} is not this rank's shard size {self.bounds[self.rank][1]}")
            as_u8 = t.dtype == torch.bool
            src = t.view(torch.uint8) if as_u8 else t
            full = self._buffer(key, src)
            src = src.contiguous()
            if self.equal and self.dst is not None:
                views = [full[s:s + c] for s, c in self.bounds] if self.rank == self.dst else None
                dist.gather(src, views, dst=self.dst, group=self.group)
            elif self.equal:
                dist.all_gather_into_tensor(full, src, group=self.group)
            else:  # ragged shards: pad every shard to the largest, gather, then compact into the global batch
                maxc = max(c for _, c in self.bounds)
                padded = torch.zeros((maxc,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
                padded[: src.shape[0]] = src
                stage = torch.empty((self.world * maxc,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
                dist.all_gather_into_tensor(stage, padded, group=self.group)
                for r, (s0, c) in enumerate(self.bounds):
                    full[s0:s0 + c] = stage[r * maxc:r * maxc + c]
            out[key] = full.view(torch.bool) if as_u8 else full
        return out

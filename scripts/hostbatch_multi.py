"""torchrun --nproc-per-node N scripts/hostbatch_multi.py : N-rank check of the single host-side batch on GPUs.  Every rank
steps its CartPole / Humanoid shard through HostBatchPipeline (both modes); rank 0 checks each consumed global batch against
a single-process run of the same global env range (seeds use the global env index, so the shards must reproduce it)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnasium_b200  # noqa: E402
from gymnasium_b200.distributed import HostBatchPipeline, env_rank_world  # noqa: E402

rank, local, world = env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
all_ok = True
for env_id, n, steps in (("CartPole-v1", 4096, 25), ("Humanoid-v5", 64, 8)):
    total = n * world
    rs = np.random.default_rng(3)
    acts = (rs.uniform(-0.4, 0.4, size=(steps, total, 17)).astype(np.float32) if env_id == "Humanoid-v5"
            else rs.integers(0, 2, size=(steps, total)))
    expect = None
    if rank == 0:
        ref = gymnasium_b200.make_vec(env_id, num_envs=total, output="numpy")
        ref.reset(seed=21)
        expect = [ref.step(a) for a in acts]
    for mode, fast in (("dma", True), ("dma", False), ("nccl", False)):
        ok = True
        env = gymnasium_b200.make_vec(env_id, num_envs=n, copy=False, out_buffers=3, env_offset=rank * n)
        env.reset(seed=21)
        pipe = HostBatchPipeline(env, world, rank, tag=f"multi_{env_id}_{int(fast)}", depth=3, mode=mode, fast=fast)
        got = []
        for k in range(steps):
            t = pipe.submit(acts[k, rank * n:(rank + 1) * n])
            if pipe.is_consumer and t >= 1:  # read the batch BEFORE releasing its slot: the other ranks run ahead
                got.append({key: v.copy() for key, v in pipe.consume(t - 1, ack=False).items()})
                pipe.release(t - 1)
        if pipe.is_consumer:
            got.append({key: v.copy() for key, v in pipe.consume(steps - 1, ack=False).items()})
            pipe.release(steps - 1)
        pipe.drain()
        dist.barrier()
        pipe.close()
        if rank == 0:
            for k in range(steps):
                for j, key in enumerate(("obs", "reward", "terminated", "truncated")):
                    if not np.array_equal(got[k][key], expect[k][j]):
                        ok = False
                        bad = np.nonzero(np.asarray(got[k][key] != expect[k][j]).reshape(total, -1).any(axis=1))[0]
                        if k < 3:
                            print(f"MISMATCH {env_id} mode={mode} fast={fast} step {k} key {key}: {len(bad)} rows, ranks "
                                  f"{sorted(set((bad // n).tolist()))}, first rows {bad[:4].tolist()}", flush=True)
            all_ok = all_ok and ok
            print(f"{env_id} world={world} mode={mode} fast={fast}: {'ok' if ok else 'FAILED'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if all_ok else 1)

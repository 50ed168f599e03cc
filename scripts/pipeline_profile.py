"""Host-side cost of the end-to-end paths at N=1 (CartPole-v1, 65536 envs): where the microseconds of one step go."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnasium_b200  # noqa: E402
from gymnasium_b200.distributed import HostBatchPipeline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env_id = sys.argv[2] if len(sys.argv) > 2 else "CartPole-v1"
rs = np.random.default_rng(0)
host = rs.integers(0, 2, size=(16, n)) if env_id == "CartPole-v1" else rs.uniform(-0.4, 0.4, size=(16, n, 17)).astype(np.float32)



def loop(fn, iters=2000):
    for k in range(50):
        fn(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(iters):
        fn(k)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6


iters = 2000 if env_id == "CartPole-v1" else 40
env = gymnasium_b200.make_vec(env_id, num_envs=n, copy=False)
env.reset(seed=0)
dev = [torch.as_tensor(h).cuda() for h in host]
print(f"step(device actions), no sync      : {loop(lambda k: env.step(dev[k % 16]), iters):8.1f} us")
print(f"step(host actions) -> torch, no sync: {loop(lambda k: env.step(host[k % 16]), iters):8.1f} us")
envn = gymnasium_b200.make_vec(env_id, num_envs=n, copy=False, output="numpy")
envn.reset(seed=0)
print(f"step(host actions) -> numpy (blocking): {loop(lambda k: envn.step(host[k % 16]), iters):8.1f} us")
envp = gymnasium_b200.make_vec(env_id, num_envs=n, copy=False, out_buffers=3)
envp.reset(seed=0)
pipe = HostBatchPipeline(envp, 1, 0, tag="prof", depth=3)


def pstep(k):
    t = pipe.submit(host[k % 16])
    if t >= 1:
        pipe.consume(t - 1)


print(f"pipeline submit + consume(k-1)        : {loop(pstep, iters):8.1f} us")
import cProfile
import pstats

pr = cProfile.Profile()
pr.enable()
for k in range(iters):
    pstep(k)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

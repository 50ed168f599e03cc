"""Where does an end-to-end step() (host numpy actions -> host numpy results) spend its time?"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gymnasium_b200

n = 65536
env = gymnasium_b200.make_vec("CartPole-v1", num_envs=n, copy=False, output="numpy")
env.reset(seed=0)
acts = np.random.default_rng(0).integers(0, 2, size=(16, n)).astype(np.int64)
for k in range(20):
    env.step(acts[k % 16])
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(500):
    env.step(acts[k % 16])
print("e2e us/step", (time.perf_counter() - t0) / 500 * 1e6)
pr = cProfile.Profile()
pr.enable()
for k in range(500):
    env.step(acts[k % 16])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
# raw copy costs
pin_a = torch.empty(n, dtype=torch.int64, pin_memory=True)
dev_a = torch.empty(n, dtype=torch.int64, device="cuda")
obs = torch.empty((n, 4), device="cuda")
pin_o = torch.empty((n, 4), pin_memory=True)
for name, fn in [("h2d 512KB", lambda: dev_a.copy_(pin_a, non_blocking=True)),
                 ("d2h 1MB", lambda: pin_o.copy_(obs, non_blocking=True))]:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
        torch.cuda.current_stream().synchronize()
    print(name, "us", (time.perf_counter() - t0) / 200 * 1e6)
t0 = time.perf_counter()
for _ in range(200):
    pin_a.copy_(torch.from_numpy(acts[0]))
print("host memcpy 512KB into pinned us", (time.perf_counter() - t0) / 200 * 1e6)

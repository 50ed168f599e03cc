"""Small Humanoid run for compute-sanitizer (racecheck / memcheck): warp-per-env kernel, a few CTAs, long enough for
contacts, joint limits, terminations and autoresets to occur."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gymnasium_b200

n = int(os.environ.get("B2E_N", "40"))
steps = int(os.environ.get("B2E_STEPS", "70"))
e = gymnasium_b200.make_vec("Humanoid-v5", num_envs=n, impl="warp", max_episode_steps=50)
e.reset(seed=3)
rs = np.random.default_rng(0)
nterm = 0
for t in range(steps):
    out = e.step(rs.uniform(-0.4, 0.4, size=(n, 17)).astype(np.float32))
    nterm += int(out[2].sum())
torch.cuda.synchronize()
print("steps", steps, "terminations", nterm, "overflow", bool(e.buffer_overflow()))

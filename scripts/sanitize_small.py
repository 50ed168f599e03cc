"""Small mixed run for compute-sanitizer memcheck / synccheck / initcheck: every family, both autoreset flavours."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gymnasium_b200

rs = np.random.default_rng(0)
for env_id, n, steps, kw in [("Blackjack-v1", 300, 30, {}), ("Blackjack-v1", 300, 30, {"autoreset_mode": "SameStep"}),
                             ("Taxi-v4", 300, 20, {}), ("FrozenLake-v1", 300, 20, {"map_name": "8x8", "autoreset_mode": "SameStep"}),
                             ("CartPole-v1", 300, 40, {}), ("Acrobot-v1", 100, 10, {}), ("Pendulum-v1", 100, 10, {}),
                             ("LunarLander-v3", 70, 90, {}), ("Humanoid-v5", 20, 35, {"impl": "warp", "autoreset_mode": "SameStep", "max_episode_steps": 20}),
                             ("Humanoid-v5", 12, 30, {"impl": "thread"})]:
    e = gymnasium_b200.make_vec(env_id, num_envs=n, **kw)
    e.reset(seed=1)
    sp = e.single_action_space
    for t in range(steps):
        if hasattr(sp, "n"):
            a = rs.integers(0, sp.n, n)
        else:
            a = rs.uniform(sp.low, sp.high, size=(n,) + sp.shape).astype(np.float32)
        out = e.step(a)
    torch.cuda.synchronize()
    print(env_id, kw, "ok", int(np.asarray(out[2].cpu() if hasattr(out[2], "cpu") else out[2]).sum()), flush=True)

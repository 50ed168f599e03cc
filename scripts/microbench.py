"""Ad-hoc device timings of the step / rollout kernels (CUDA events); used during development, not the bench contract."""
import ctypes as C
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gymnasium_b200


def time_calls(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    res = {}
    for fam, nacts, kw in [("CartPole-v1", 2, {}), ("FrozenLake-v1", 4, {"map_name": "8x8"})]:
        for n in [65536, 1 << 20, 1 << 24]:
            for adt in [torch.int64, torch.uint8]:
                env = gymnasium_b200.make_vec(fam, num_envs=n, copy=False, **kw)
                env.reset(seed=0)
                acts = torch.randint(0, nacts, (n,), device="cuda").to(adt)
                t = time_calls(lambda: env.step(acts), 200)
                res[f"{fam} step n={n} act={str(adt)[6:]}"] = dict(us=t * 1e6, steps_per_s=n / t)
                # graph replay of 16 steps
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    env.step(acts)
                    torch.cuda.synchronize()
                    with torch.cuda.graph(g, stream=s):
                        for _ in range(16):
                            env.step(acts)
                t = time_calls(g.replay, 50) / 16
                res[f"{fam} graph16 n={n} act={str(adt)[6:]}"] = dict(us=t * 1e6, steps_per_s=n / t)
                del env, g
            for rng in ["numpy", "philox"]:
                env = gymnasium_b200.make_vec(fam, num_envs=n, rng=rng, **kw)
                env.reset(seed=0)
                K = 64 if n <= (1 << 20) else 8
                t = time_calls(lambda: env.rollout(K), 10, warm=2)
                res[f"{fam} rollout K={K} n={n} rng={rng}"] = dict(us_per_step=t / K * 1e6, steps_per_s=n * K / t)
                del env
            torch.cuda.empty_cache()
    for k, v in res.items():
        print(k, json.dumps(v))


if __name__ == "__main__":
    main()

"""Sparse lane mapping sweep for the divergent physics kernels + FrozenLake grid check (device timings, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gymnasium_b200

def timed(fn, iters, warm):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3

n = 16384
for lanes in [int(x) for x in os.environ.get("B2E_LANES", "32,16,8,4,2").split(",")]:
    print("grouping", os.environ.get("B2E_GROUPING", "0"), end=" ")
    e = gymnasium_b200.make_vec("LunarLander-v3", num_envs=n, copy=False)
    e._cfg.lanes_per_warp = lanes
    if hasattr(e._cfg, "grouping"):
        e._cfg.grouping = int(os.environ.get("B2E_GROUPING", "0"))
    e.reset(seed=0)
    a = torch.randint(0, 4, (8, n), device="cuda")
    k = [0]
    def f():
        e.step(a[k[0] % 8]); k[0] += 1
    t = timed(f, 150, 150)
    print(f"LunarLander n={n} lanes/warp {lanes:2d}: {t*1e6:8.1f} us/step  {n/t:.3e} steps/s")
if os.environ.get("B2E_LANDER_ONLY"):
    sys.exit(0)
n = 8192
for lanes in [32, 16, 8, 4, 2]:
    e = gymnasium_b200.make_vec("Humanoid-v5", num_envs=n, copy=False, impl="thread")
    e._cfg.lanes_per_warp = lanes
    e.reset(seed=0)
    a = (torch.rand((4, n, 17), device="cuda") * 0.8 - 0.4).float()
    k = [0]
    def f():
        e.step(a[k[0] % 4]); k[0] += 1
    t = timed(f, 12, 14)
    print(f"Humanoid n={n} lanes/warp {lanes:2d}: {t*1e3:8.2f} ms/step  {n/t:.3e} steps/s")
nfl = 1 << 20
fls = [gymnasium_b200.make_vec("FrozenLake-v1", num_envs=nfl, map_name="8x8", copy=False, env_offset=j * nfl) for j in range(4)]
for f_ in fls: f_.reset(seed=0)
a = torch.randint(0, 4, (nfl,), device="cuda")
k = [0]
def f():
    fls[k[0] % 4].step(a); k[0] += 1
t = timed(f, 80, 12)
print(f"FrozenLake 1M ring: {t*1e6:.2f} us/step {nfl/t:.3e} steps/s {98*nfl/t/1e9:.0f} GB/s")

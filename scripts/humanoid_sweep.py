"""One-process sweep of the Humanoid kernel mappings / launch shapes (tuning aid; never a bench number).
Every configuration replays the same seeded batch from the same steady-state start and must produce identical outputs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gymnasium_b200

dev = torch.device("cuda:0")
n = int(os.environ.get("B2E_N", "8192"))
steps = int(os.environ.get("B2E_STEPS", "6"))
g = torch.Generator(device=dev).manual_seed(1)
acts = [torch.rand((n, 17), device=dev, generator=g) * 0.8 - 0.4 for _ in range(60 + steps)]
# (impl, envs per CTA, schedule bits: 1 = no CTA barrier, 2 = no work grouping)
configs = [("thread", 0, 0)] + [("warp", w, sch) for w, sch in ((1, 0), (4, 1), (2, 0), (4, 2), (4, 0), (8, 2), (8, 0))]
if len(sys.argv) > 1:
    configs = [(c.split(":")[0], int(c.split(":")[1]), int(c.split(":")[2])) for c in sys.argv[1:]]
ref = None
for impl, per_cta, schedule in configs:
    e = gymnasium_b200.make_vec("Humanoid-v5", num_envs=n, copy=False, impl=impl)
    e._cfg.envs_per_cta, e._cfg.schedule = per_cta, schedule
    e.reset(seed=0)
    for t in range(60):
        e.step(acts[t])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(steps):
        out = e.step(acts[60 + t])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    sig = (out[0].double().sum().item(), out[1].sum().item(), int(out[2].sum().item()))
    same = "" if ref is None else ("same" if sig == ref else f"DIFFERENT {sig} vs {ref}")
    ref = ref or sig
    barrier = "n" if (schedule & 9 or per_cta == 1) else ("per-step" if schedule & 4 else "y")
    grouped = "n" if (per_cta == 1 or schedule & 2 or (schedule & 1 and not schedule & 8)) else "y"
    label = f"warp envs/CTA={per_cta} barrier={barrier} grouped={grouped}" if impl == "warp" else "thread"
    print(f"Humanoid n={n} {label:46s}: {dt*1e3:8.2f} ms/step  {n/dt:.3e} steps/s  {same}", flush=True)
    del e

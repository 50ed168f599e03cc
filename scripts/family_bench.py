"""Device throughput of every family's step kernel at a DRAM-sized batch (CUDA events; inputs > L2)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gymnasium_b200

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
# algorithmic bytes per env-step: state r+w, ctrl r+w, action, obs, reward f64, flags (+ rng for the tabular families)
FAM = [
    ("CartPole-v1", {}, 2, None, 106),
    ("MountainCar-v0", {}, 3, None, 2 * 16 + 8 + 8 + 8 + 8 + 2),
    ("MountainCarContinuous-v0", {}, None, 1.0, 2 * 16 + 2 + 8 + 4 + 8 + 8 + 2),
    ("Pendulum-v1", {}, None, 2.0, 2 * 16 + 8 + 4 + 12 + 8 + 2),
    ("Acrobot-v1", {}, 3, None, 2 * 32 + 8 + 8 + 24 + 8 + 2),
    ("FrozenLake-v1", {"map_name": "8x8"}, 4, None, 98),
    ("CliffWalking-v1", {}, 4, None, 98),
    ("Taxi-v4", {}, 6, None, 98),
    ("Blackjack-v1", {}, 2, None, 118),
]
n = 1 << 23
for fam, kw, nact, scale, nbytes in FAM:
    env = gymnasium_b200.make_vec(fam, num_envs=n, copy=False, **kw)
    env.reset(seed=0)
    if nact:
        acts = [torch.randint(0, nact, (n,), device="cuda") for _ in range(4)]
    else:
        acts = [((torch.rand((n, 1), device="cuda") * 2 - 1) * scale).float() for _ in range(4)]
    for k in range(5):
        env.step(acts[k % 4])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(20):
        env.step(acts[k % 4])
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 20 * 1e-3
    bw = nbytes * n / t / 1e9
    print(json.dumps({"family": fam, "n": n, "us_per_step": t * 1e6, "steps_per_s": n / t, "algorithmic_bytes": nbytes,
                      "GBs": bw, "frac_of_hbm_peak": bw / PEAK}))
    del env, acts
    torch.cuda.empty_cache()

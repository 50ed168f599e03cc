"""CartPole step kernel: CTA size sweep at N=65536 (graph-chained launches, L2-resident and HBM-cold ring)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gymnasium_b200

def chain(envs, acts, L=60, reps=40):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for e in envs: e.step(acts)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for k in range(L):
                envs[k % len(envs)].step(acts)
    for _ in range(30): g.replay()   # also walks every batch into its steady state (~5 % of the lanes on an autoreset call)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * L)

n = 65536
acts = torch.randint(0, 2, (n,), device="cuda")
ring = [gymnasium_b200.make_vec("CartPole-v1", num_envs=n, copy=False, env_offset=j * n) for j in range(30)]
for e in ring: e.reset(seed=0)
print("speculative RNG load:", "on" if os.environ.get("B2E_CARTPOLE_SPEC_RNG") else "off", " PDL:", "on" if os.environ.get("B2E_PDL") else "off")
for blk in [int(x) for x in os.environ.get('B2E_SWEEP_BLOCKS', '32,64,96,128,192,256,448,1024').split(',')]:
    for e in ring: e._cfg.step_block = blk
    print(f"block {blk:5d}: L2-resident {chain(ring[:1], acts):.3f} us   HBM-cold ring {chain(ring, acts):.3f} us")

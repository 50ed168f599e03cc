"""Turns the raw gpurun_out/<tag>/ captures into the committed evidence under profiles/ (run here, no GPU needed)."""
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import summarise

tag, rnd = sys.argv[1], sys.argv[2]  # e.g. r1b r1
src = os.path.join("gpurun_out", tag)
os.makedirs("profiles", exist_ok=True)
KEYS = {"step": "cartpole_step_kernel<int64>", "big": "cartpole_step_kernel<int64> N=16M",
        "rollout": "cartpole_rollout_kernel<philox>", "lake": "frozenlake_step_kernel<int64>",
        "lander": "lunarlander_step_kernel<int64>", "humanoid": "humanoid_step_warp_kernel<float, 10>",
        "land": "land_outputs_kernel (CartPole-v1 65536 envs: 1.7 MB of step outputs into the mapped host batch)"}
summary, lines = {}, []
for name, key in KEYS.items():
    p = os.path.join(src, f"ncu_{name}.ncu-rep")
    if not os.path.exists(p):
        continue
    units, rows = summarise(p)
    if not rows:
        continue

    def val(r, k, scale_unit=None):
        v = r.get(k)
        return v

    def to_bytes(v, unit):
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

    rd = sum(to_bytes(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) for r in rows) / len(rows)
    wr = sum(to_bytes(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"]) for r in rows) / len(rows)
    dur = sum(r["gpu__time_duration.sum"] for r in rows) / len(rows)
    if units["gpu__time_duration.sum"] == "ms":
        dur *= 1e3
    r0 = rows[0]
    summary[key] = {
        "launches_captured": len(rows), "duration_us_cold": dur, "dram_read_bytes_per_launch": rd,
        "dram_write_bytes_per_launch": wr, "dram_bytes_per_launch": rd + wr,
        "dram_throughput_pct": r0.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_throughput_pct": r0.get("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": r0.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": r0.get("launch__registers_per_thread"), "grid": r0.get("launch__grid_size"),
        "block": r0.get("launch__block_size"),
        "fp64_pipe_pct": r0.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": r0.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions": r0.get("smsp__inst_executed.sum"),
        "kernel_name": r0.get("Kernel Name"),
        "threads_active_per_instruction": r0.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "tensor_pipe_cycles_active_pct": r0.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_instructions": r0.get("sm__inst_executed_pipe_tensor.sum"),
        "fp64_pipe_instructions": r0.get("sm__inst_executed_pipe_fp64.sum"),
        "fma_pipe_pct": r0.get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "local_load_sectors": r0.get("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
        "local_store_sectors": r0.get("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
        "l1_hit_rate_pct": r0.get("l1tex__t_sector_hit_rate.pct"),
        "shared_mem_per_block_dynamic": r0.get("launch__shared_mem_per_block_dynamic"),
        "stall_cycles_per_issue": {k: r0.get(f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio")
                                   for k in ("long_scoreboard", "barrier", "wait", "no_instruction", "short_scoreboard",
                                             "branch_resolving", "math_pipe_throttle")},
    }
# bench.py reads profiles/ncu_summary.json for roofline.traffic
flat = {k.replace(" N=16M", "_16M"): v for k, v in summary.items()}
json.dump(flat, open("profiles/ncu_summary.json", "w"), indent=1)
json.dump(flat, open(f"profiles/{rnd}_ncu_summary.json", "w"), indent=1)
# launch list of the bench command: share of each kernel in the captured launches
lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = {}
    for r in rows[hdr + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        name = r[ki].split("(")[0][-60:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(f"profiles/{rnd}_launches_bench.md", "w") as f:
        f.write(f"# ncu launch list of `python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-extras --e2e-steps 20` ({tag})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 2500` (the first 100 launches are skipped: ring allocation, seeding); per-launch times are cold-cache and "
                "serialised, so only the SHARES are meaningful.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |\n")
print('wrote profiles/', sorted(flat))

"""Summarise .ncu-rep captures (run here, no GPU needed): prints and optionally writes the key per-launch metrics."""
import csv
import json
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sector_hit_rate.pct"] + [
    f"smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio"
    for k in ("barrier", "wait", "no_instruction", "short_scoreboard", "branch_resolving", "math_pipe_throttle")]


def summarise(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    units = {w: rows[1][i] for w, i in idx}
    out = []
    for r in rows[2:]:
        d = {}
        for w, i in idx:
            v = r[i]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                pass
            d[w] = v
        out.append(d)
    return units, out


if __name__ == "__main__":
    res = {}
    for p in sys.argv[1:]:
        units, rows = summarise(p)
        print("==", p)
        print("units", json.dumps(units))
        for r in rows:
            print(json.dumps(r))

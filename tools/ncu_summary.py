"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of counters the roofline discussion uses.
usage: python tools/ncu_summary.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__ops_path_tensor_op_utcimma_src_int8.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utcimma_src_int8.sum.per_second", "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel", "metric", "value", "unit"])
    for n, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "?")[:90]
        for k in KEYS:
            hit = [h for h in hdr if h.endswith(k)]
            if hit:
                w.writerow([n, name, k, d[hit[0]], units[hdr.index(hit[0])]])
print("wrote", out)

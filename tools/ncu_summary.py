#!/usr/bin/env python
"""Reads a `ncu --set full` report here (no GPU needed) and prints the metrics the roofline needs, per launch.

    python tools/ncu_summary.py gpurun_out/warp_full.ncu-rep > profiles/rNN_ncu_warp.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# %s: %d profiled launch(es); ncu --set full --clock-control none (cold caches, replayed passes)" % (rep, len(data)))
    names = [d[idx["Kernel Name"]] for d in data]
    for i, n in enumerate(names):
        print("# launch %d: %s  grid %s block %s" % (i, n[:90], data[i][idx["Grid Size"]], data[i][idx["Block Size"]]))
    for m in METRICS:
        if m in idx:
            print("%-82s %-16s %s" % (m, units[idx[m]], "  ".join(d[idx[m]] for d in data)))
    if "dram__bytes_read.sum" in idx:
        def val(d, m):
            u = units[idx[m]].lower()
            mult = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
            return float(d[idx[m]].replace(",", "")) * mult
        print("%-82s %-16s %s" % ("traffic = dram__bytes_read.sum + dram__bytes_write.sum", "byte",
                                  "  ".join("%.0f" % (val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")) for d in data)))


if __name__ == "__main__":
    main()

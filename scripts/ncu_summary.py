"""Condenses an .ncu-rep (read here, on the CPU box, with `ncu -i`) into the few counters DESIGN.md and
bench.py's roofline cite.  Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [kernel-substring] > profiles/x.txt"""
import csv
import io
import subprocess
import sys

EXACT = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "sm__cycles_elapsed.max.per_second",
    "launch__grid_size",
    "launch__block_size",
    "launch__cluster_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct",
    "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
]
SUBSTR = ["tensor", "tmem", "utc"]


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print("# source: %s (ncu --set full --clock-control none; per-launch, cold-cache, serialised)" % rep)
    for n, r in enumerate(data):
        if pat and pat not in r[ki]:
            continue
        print("\n## launch %d: %s" % (n, r[ki][:160]))
        for i, h in enumerate(hdr):
            if h in EXACT or (any(s in h.lower() for s in SUBSTR) and h.endswith(("pct_of_peak_sustained_active", "pct_of_peak_sustained_elapsed", ".sum"))
                              and r[i] not in ("0", "", "n/a")):
                print("%-90s %14s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main()

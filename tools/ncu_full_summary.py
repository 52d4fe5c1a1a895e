#!/usr/bin/env python
"""Key metrics of every kernel in `ncu --set full` reports (read here, on the CPU box, with `ncu -i`):

    python tools/ncu_full_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/rNN_ncu_full_summaries.txt
"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "regs/thread"),
        ("launch__block_size", "block"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.per_cycle_active", "warps/SM"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue-active %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("smsp__inst_executed.sum", "warp instrs"),
        ("sass__inst_executed_local_loads", "local loads"), ("sass__inst_executed_local_stores", "local stores"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio")]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print("== %s" % path.split("/")[-1])
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print("  kernel %s" % name[:110])
            for key, label in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    print("    %-22s %s %s" % (label, r[i], units[i]))
        print()


if __name__ == "__main__":
    main()

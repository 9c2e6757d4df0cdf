"""Prints the handful of ncu metrics we track from a .ncu-rep (run here; no GPU needed)."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "lts__t_bytes.sum ", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum ", "sm__cycles_elapsed.max ", "launch__registers_per_thread ", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled_long_scoreboard ", "smsp__pcsamp_warps_issue_stalled_barrier ",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard ", "smsp__pcsamp_warps_issue_stalled_wait ", "smsp__pcsamp_warps_issue_stalled_branch_resolving ",
        "smsp__pcsamp_warps_issue_stalled_no_instruction ", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle ", "smsp__pcsamp_warps_issue_stalled_membar ",
        "smsp__pcsamp_sample_count", "launch__grid_size", "dram__throughput", "smsp__pcsamp_warps_issue_stalled_selected ", "smsp__pcsamp_warps_issue_stalled_sleeping ",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle ", "smsp__pcsamp_warps_issue_stalled_lg_throttle ", "smsp__pcsamp_warps_issue_stalled_tex_throttle ", "smsp__pcsamp_warps_issue_stalled_dispatch_stall ", "smsp__pcsamp_warps_issue_stalled_not_selected "]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("==", rep)
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
        print("  kernel:", name[:80])
        for h, u, v in zip(hdr, units, vals):
            if any((h + " ").startswith(k) or h == k.strip() for k in KEYS):
                print(f"    {h:75s} {v:>16s} {u}")

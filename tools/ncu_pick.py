"""Print a fixed short list of metrics from `ncu -i rep --page raw --csv` output, one column per profiled launch.
usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_pick.py [extra_metric_substring ...]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
PICK = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "launch__registers_per_thread"]
extra = sys.argv[1:]
print("kernels:", [r[hdr.index("Kernel Name")][:28] for r in data])
for i, h in enumerate(hdr):
    if h in PICK or any(e in h for e in extra):
        print(f"{h} [{units[i]}]:", [r[i] for r in data])

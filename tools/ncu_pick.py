"""Print selected metrics of an `ncu --page raw --csv` export, one column per captured launch.
    python tools/ncu_pick.py gpurun_out/r02_ncu_full_attn_tc.csv [substring ...]"""
import csv
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fmaheavy", "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct",
           "launch__registers_per_thread ", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct",
           "smsp__average_warp", "smsp__cycles_active.avg ", "sm__cycles_elapsed.avg ", "smsp__inst_executed.sum ", "sm__pipe_tensor"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pick = sys.argv[2:] or DEFAULT
for i, h in enumerate(hdr):
    if any(p.strip() in h for p in pick):
        print(f"{h} [{units[i]}]: " + " | ".join(r[i][:28] for r in data))

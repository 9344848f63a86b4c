"""Selected raw metrics of one `ncu --set full` capture → a small CSV for profiles/.
Usage: ncu -i capture.ncu-rep --page raw --csv > raw.csv; python tools/ncu_select.py raw.csv [launch index] > profiles/<name>.csv"""
import csv, sys
PREFIX = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "dram__throughput.avg.pct",
          "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size",
          "launch__grid_size", "launch__occupancy_limit", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
          "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__throughput.avg.pct",
          "sm__inst_executed_pipe_fma", "sm__pipe_fma", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct",
          "sm__issue_active.avg.pct", "sm__warps_active.avg.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
          "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled", "sm__sass_thread_inst_executed_op_f",
          "smsp__sass_thread_inst_executed_op_f", "sm__inst_executed.avg.per_cycle_active", "smsp__cycles_active.avg")
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr, units, vals = rows[0], rows[1], rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit", "value"])
w.writerow(["kernel", "", vals[hdr.index("Kernel Name")]])
for h, u, v in sorted(zip(hdr, units, vals)):
    if h.startswith(PREFIX):
        w.writerow([h, u, v])

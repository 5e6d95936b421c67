"""Writes the judged summary of an .ncu-rep (raw page) as `metric,unit,value` rows:  python tools/ncu_profile_csv.py rep.ncu-rep "comment" > profiles/x.csv"""
import csv
import subprocess
import sys

KEEP = ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__throughput.avg.pct", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "lts__throughput.avg.pct",
        "sm__pipe_tensor_cycles_active.avg.pct", "sm__throughput.avg.pct", "sm__warps_active.avg.pct",
        "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma",
        "sm__inst_executed_pipe_lsu", "sm__pipe_alu_cycles_active", "sm__pipe_fma_cycles_active", "sm__pipe_fmaheavy",
        "sm__inst_executed_pipe_tensor")
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print("# " + (sys.argv[2] if len(sys.argv) > 2 else rep))
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in KEEP) and "per_second" not in h.replace("dram__bytes_write.sum.per_second", ""):
            if "_ratio" in h and "issue_stalled" in h and "per_issue_active" not in h:
                continue
            print(f"{h},{units[i]},{r[i]}")

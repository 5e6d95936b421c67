import csv,sys,subprocess,collections
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
def show(pat):
    for i,h in enumerate(hdr):
        if any(p in h for p in pat):
            print(f"{h} [{units[i]}] = {[r[i] for r in rows[2:]]}")
show(["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct","sm__throughput.avg.pct","sm__warps_active.avg.pct","launch__registers","launch__occupancy_limit","launch__waves","smsp__inst_executed.sum ","smsp__issue_active.avg.pct","smsp__thread_inst_executed_per_inst","bank_conflicts_pipe_lsu_mem_shared.sum","launch__grid_size","smsp__inst_executed.sum"])
print("--- stall reasons (pct of warp active or ratio)")
for i,h in enumerate(hdr):
    if "warp_issue_stalled" in h and ("per_warp_active" in h or "ratio" in h) :
        v=rows[2][i]
        try:
            if float(v)>1.0: print(f"   {h} = {v}")
        except: pass

"""Summarise an .ncu-rep: python tools/ncu_read.py file.ncu-rep [source]"""
import csv, subprocess, sys, io
f = sys.argv[1]
out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
keys = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "smsp__warp_issue_stalled", "sm__throughput.avg.pct"]
for r in rows[2:]:
    print("-----")
    for i, h in enumerate(hdr):
        if any(h.startswith(k) or k in h for k in keys):
            if "stalled" in h and "pct" not in h and "ratio" not in h: continue
            print(f"  {h} = {r[i]}")

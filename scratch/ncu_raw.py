import csv, subprocess, sys
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = sys.argv[2:] or ["gpu__time_duration.sum","launch__grid_size","launch__block_size","launch__cluster","launch__registers_per_thread","smsp__inst_executed.sum","sm__cycles_elapsed.max","smsp__cycles_active.avg","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_bytes.sum","lts__t_sectors_op_read.sum","lts__t_sectors_op_write.sum","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct","smsp__pcsamp_warps_issue_stalled","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct","l1tex__t_bytes","smsp__average_warps_issue_stalled"]
for h,u,v in zip(hdr,units,vals):
    if any(h.startswith(w) for w in want) and not h.endswith("_not_issued"):
        print(f"{h:80s} {u:16s} {v}")

"""Key metrics of an ncu report for the latency-bound imprint kernel: python scratch/ncu_key.py REPORT [N_IMPRINTS]."""
import csv, subprocess, sys
rep = sys.argv[1]; nimp = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
m = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
def g(k):
    try: return float(m[k].replace(",", ""))
    except Exception: return float("nan")
dur = g("gpu__time_duration.sum"); unit = u.get("gpu__time_duration.sum")
dur_us = dur * {"ms": 1e3, "us": 1.0, "s": 1e6, "ns": 1e-3}.get(unit, 1.0)
warps = g("launch__grid_size") * g("launch__block_size") / 32
ins = g("smsp__inst_executed.sum")
print(f"duration {dur_us:.1f} us = {dur_us/nimp:.2f} us/imprint; grid {g('launch__grid_size'):.0f} x {g('launch__block_size'):.0f}, regs {g('launch__registers_per_thread'):.0f}")
print(f"warp-instructions {ins:.3g} = {ins/warps/nimp:.0f} per warp per imprint; issue active {g('smsp__issue_active.avg.per_cycle_active'):.3f}/cycle/scheduler")
ldr, lds = g("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), g("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
str_, sts = g("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"), g("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum")
print(f"global ld requests {ldr:.3g} ({ldr/warps/nimp:.1f}/warp/imprint), sectors/request {lds/ldr:.2f}; st requests {str_:.3g}, sectors/request {sts/str_:.2f}")
print(f"local ld instr {g('sass__inst_executed_local_loads'):.3g}, local st instr {g('sass__inst_executed_local_stores'):.3g}")
print(f"L2 sectors {g('lts__t_sectors.sum'):.3g}, L2 hit {g('lts__t_sector_hit_rate.pct'):.1f} %, lts throughput {g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} %, l1tex throughput max {g('l1tex__throughput.max.pct_of_peak_sustained_elapsed'):.1f} %")
print("stall cycles per issued instruction:")
for k in sorted(m):
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
        v = g(k)
        if v >= 0.15: print(f"   {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:.2f}")

"""Key metrics of one ncu report (first kernel): python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(f"ncu -i {sys.argv[1]} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, vals = rows[0], rows[2]
d = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for k in keys:
    for h in hdr:
        if h.startswith(k):
            print(f"{h:75s} {d[h]}")
print("-- stall cycles per issued instruction --")
st = [(float(d[h]), h.split("issue_stalled_")[-1].replace("_per_issue_active.ratio", "")) for h in hdr
      if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h and d[h] not in ("", "n/a")]
for v, n in sorted(st, reverse=True)[:8]:
    print(f"{n:30s} {v:8.3f}")

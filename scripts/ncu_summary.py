"""Text summary of one kernel of an .ncu-rep (ncu -i ... --page raw --csv): the counters the
roofline discussion in DESIGN.md uses.  usage: python scripts/ncu_summary.py rep.ncu-rep "header" > out.txt"""
import csv
import subprocess
import sys

rep, header = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
print(header)
for r in rows[2:]:
    for kname in keys:
        if kname in hdr:
            i = hdr.index(kname)
            print("%-75s %s %s" % (kname, r[i], units[i]))
    print("# warp stall reasons (warps per issue-active cycle)")
    stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
              and r[i] not in ("", "nan", "-nan")]
    for v, h in sorted(stalls, reverse=True)[:9]:
        print("%-75s %.3f" % (h, v))

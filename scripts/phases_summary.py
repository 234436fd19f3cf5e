"""ncu `--page raw --csv` capture of scripts/phase_ncu.py (one row per launch, one column per metric) -> metric x launch table.
    python scripts/phases_summary.py gpurun_out/r02_phases_raw.csv > profiles/<round>_phases_table.txt"""
import csv, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_sample_count"]
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["launch %s" % r[0] for r in data]
print("%-78s" % "metric" + "".join("%16s" % n for n in names))
for k in KEEP:
  if k in hdr:
    i = hdr.index(k)
    print("%-78s" % ("%s [%s]" % (k, units[i])) + "".join("%16s" % r[i] for r in data))

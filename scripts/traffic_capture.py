"""Profiling aid: runs the steady-state window (and the cold window) of the headline workload as their own launches, for
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:pdhg_coop ... python scripts/traffic_capture.py
Launch order of pdhg_coop_kernel: [tables, spin-up march, tables, STEADY window (index 3)], [tables, warm-up march, tables, COLD window (index 7)].
`scripts/traffic_from_ncu.py` turns the CSV into profiles/traffic_<workload>[_cold].json stamped with the kernel-source hash."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
pb = bench.make_problem("cfg3_tsp65")
r = bench.run_ours_block(pb, iters, 3, 0, spinup=600)
print("steady: iters", r["iters"], "inner/iter", r["n_inner"] / r["iters"], "kernel_ms", r["kernel_ms"], flush=True)
r = bench.run_ours_block(pb, 20, 5, 0, spinup=0)
print("cold: iters", r["iters"], "inner/iter", r["n_inner"] / r["iters"], "kernel_ms", r["kernel_ms"], flush=True)

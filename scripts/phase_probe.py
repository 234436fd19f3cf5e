"""Scratch: per-phase device time of the cooperative kernel on the bench workloads."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
for name, its in (("cfg3_tsp65", int(sys.argv[1]) if len(sys.argv) > 1 else 100), ("cfg3_tsp2", 2000)):
  pb = bench.make_problem(name)
  r = bench.run_ours_block(pb, its, 3, 0)
  pt = r["solver"].phase_times_ms()
  print(name, "iters", r["iters"], "inner/iter %.2f" % (r["n_inner"] / r["iters"]), "ms/iter %.4f" % (r["kernel_ms"] / r["iters"]),
        {k: round(v / r["iters"] * 1e3, 1) for k, v in pt.items()}, "us/iter per phase")

"""Scratch: per-phase device time of the cooperative kernel on the bench workloads.
usage: phase_probe.py [iters] [libvariant ...]   (PDHG_LIB env selects the .so)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pdhg-optimal-control_b200"))
from pdhg_b200 import _lib
import bench
its = int(sys.argv[1]) if len(sys.argv) > 1 else 100
spin = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for name, n in (("cfg3_tsp65", its), ("cfg3_tsp2", 2000)):
  pb = bench.make_problem(name)
  r = bench.run_ours_block(pb, n, 3, 0, spinup=(600 if name == "cfg3_tsp65" else 0))
  pt = r["solver"].phase_times_ms()
  print(os.path.basename(_lib.LIB_PATH), name, "iters", r["iters"], "inner/iter %.2f" % (r["n_inner"] / r["iters"]), "us/iter %.1f" % (r["kernel_ms"] / r["iters"] * 1e3),
        {k[:9]: round(v / r["iters"] * 1e3, 1) for k, v in pt.items()}, flush=True)

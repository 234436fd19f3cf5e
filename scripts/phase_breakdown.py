"""Diagnostic: per-phase microseconds per outer iteration of a workload of bench.py (PDHG_PROFILE=1 in-kernel timers, CTA 0).
  PDHG_PROFILE=1 python scripts/phase_breakdown.py cfg3_tsp2 2000 0"""
import json, os, sys
os.environ.setdefault("PDHG_PROFILE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_tsp2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
spin = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pb = bench.make_problem(name)
r = bench.run_ours_block(pb, iters, 3, 0, spinup=spin)
ph = {k: round(v / max(r["iters"], 1) * 1e3, 2) for k, v in r["solver"].phase_times_ms().items()}
print(json.dumps({"workload": name, "iters": r["iters"], "us_per_iter": r["ms"] / r["iters"] * 1e3, "inner_sweeps_per_iter": r["n_inner"] / r["iters"],
                  "phase_us_per_iter": ph}))

"""Profiling aid: after a short march that fills the workspace with realistic data, launch each phase of the cooperative
kernel as its own kernel launch (pdhg_debug_phase), for `ncu -k regex:pdhg_coop -s <skip> ...`.
Launch order after the march: A, B pass 1, B pass 2, B pass 3, C, D (+reduction), D with 2 and with 5 sweeps fused."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_tsp65"
pb = bench.make_problem(name)
r = bench.run_ours_block(pb, 20, 3, 0, spinup=int(sys.argv[2]) if len(sys.argv) > 2 else 600)
s = r["solver"]
print("march launches so far:", s.launch_count, flush=True)
for ph, mask in ((0, 7), (1, 1), (1, 2), (1, 4), (2, 7), (3, 7), (3, 2), (3, 5)):     # (3, n <= 5): n dual sweeps fused in one pass
  s.debug_phase(ph, mask, pb["stepsz"] * 1.5, 1)
print("phase launches done:", s.launch_count)

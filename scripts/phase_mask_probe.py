"""Scratch: time single phases of the cooperative kernel with parts switched off (pdhg_debug_phase pass masks)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pb = bench.make_problem("cfg3_tsp65")
r = bench.run_ours_block(pb, 20, 3, 0, spinup=600)
s = r["solver"]
import torch
def t(ph, mask, reps=50):
  s.debug_phase(ph, mask, pb["stepsz"] * 1.5, 3)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  s.debug_phase(ph, mask, pb["stepsz"] * 1.5, reps)
  return (time.perf_counter() - t0) / reps * 1e6
for ph, masks in ((0, (7, 1, 2, 4, 3, 5, 6)), (1, (1, 2, 4, 7)), (2, (7,)), (3, (7,))):
  print("phase", ph, {m: round(t(ph, m), 1) for m in masks}, flush=True)

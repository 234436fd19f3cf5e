"""Scratch: per-sub-step cycles of the single-CTA 1-D kernel."""
import os, sys
os.environ["PDHG_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import numpy as np
for name, n in (("cfg1", 3000), ("cfg2", 5000)):
  pb = bench.make_problem(name)
  r = bench.run_ours_block(pb, n, 3, 0)
  out = np.zeros(16)
  from pdhg_b200 import _lib
  _lib._check(r["solver"].lib.pdhg_phase_times(r["solver"]._h, _lib._hptr(out)))
  names = ("residual", "fft", "solve", "ifft", "phi_upd", "dual", "decide", "loop")
  print(name, "iters", r["iters"], "us/iter %.2f" % (r["kernel_ms"] * 1e3 / r["iters"]), "inner/iter %.2f" % (r["n_inner"] / r["iters"]),
        {k: int(v / r["iters"]) for k, v in zip(names, out)}, "cycles/iter total", int(out[:8].sum() / r["iters"]))

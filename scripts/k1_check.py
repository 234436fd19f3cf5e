"""Scratch: per-instance iteration counts of the K=1 Green kernel vs the general FFT kernel vs the oracle (cfg4-like instances)."""
import os, sys, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pdhg-optimal-control_b200"))
import numpy as np
from pdhg_b200 import run_example as rx, set_fns
from pdhg_b200.update_fns_in_pdhg import clear_handles
B, nx, nblk = 24, 1024, 4
rng = np.random.default_rng(0)
A_, th, u = rng.uniform(0.5, 1.5, 4096)[:B], rng.uniform(0, 2 * np.pi, 4096)[:B], rng.uniform(0, 1, 4096)[:B]
x_arr = rx.make_x_arr(1, nx, 1, 2.0, 2.0)
g = A_[:, None] * np.sin(np.pi * x_arr[0, :, 0][None, :] + th[:, None])
with contextlib.redirect_stdout(io.StringIO()):
  fns = set_fns.set_up_example_fns(1, 1, 0)
T = nblk / 256.0
out = {}
for mode in ("k1", "fft"):
  clear_handles()
  if mode == "fft": os.environ["PDHG_NO_K1"] = "1"
  else: os.environ.pop("PDHG_NO_K1", None)
  phi, rho, alp, logs = rx.solve_HJ_batch(1, 1, 1, 0.002 * u, fns, nx, 1, nblk + 1, 2.0, 2.0, T, x_arr, g, 70.0, 2, 0.1, 1000000, 10000, 1e-6, 0)
  out[mode] = (phi, rho, logs.iters.copy())
d = out["k1"][2] - out["fft"][2]
print("iters k1 vs fft: max abs diff", np.abs(d).max(), "instances differing", int((np.abs(d).sum(axis=1) > 0).sum()), "of", B)
bad = np.nonzero(np.abs(d).sum(axis=1) > 0)[0].tolist()
print("differing instances", bad, [(out["k1"][2][b].tolist(), out["fft"][2][b].tolist()) for b in bad])
print("phi rel diff", np.abs(out["k1"][0] - out["fft"][0]).max() / np.abs(out["fft"][0]).max())
if len(sys.argv) > 1:
  from oracle import pdhg_numpy as orc
  fo = orc.set_up_example_fns(1, 1, 0)
  for b in (bad if sys.argv[1] == "bad" else [int(a) for a in sys.argv[1:]]):
    info = {}
    res, _ = orc.solve_HJ(1, 1, 1, float(0.002 * u[b]), fo, nx, 1, nblk + 1, 2.0, 2.0, T, x_arr, 70.0, 2, 0.1, 1000000, 10000, 1e-6, 0, g=g[b:b + 1], info=info)
    print("inst", b, "oracle", info["block_iters"], "k1", out["k1"][2][b].tolist(), "fft", out["fft"][2][b].tolist(),
          "phi rel (k1, fft)", np.abs(out["k1"][0][b] - res[0][1]).max() / np.abs(res[0][1]).max(), np.abs(out["fft"][0][b] - res[0][1]).max() / np.abs(res[0][1]).max())

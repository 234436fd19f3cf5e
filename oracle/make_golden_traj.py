"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the closed-loop trajectory simulation, produced by the reference's OWN source:
the three functions compute_traj_1d / extend_bdry_2d / compute_traj_2d are cut out of /root/reference/jaxsrc/run_example.py
(the module itself cannot be imported: it pulls in tensorflow / matplotlib) and executed unmodified with `jnp` = NumPy; the
Gaussian increments are numpy.random's after np.random.seed(seed), recorded so that the oracle and the CUDA kernel get the same
numbers.  Also checks oracle/traj_numpy.py against them.    python oracle/make_golden_traj.py
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import jax_shim  # noqa: E402
jax_shim.install()
import update_fns_in_pdhg as ref_upd  # noqa: E402  (reference, via the shim)
import set_fns as ref_set             # noqa: E402
from oracle import pdhg_numpy as orc, traj_numpy as otr  # noqa: E402
from oracle.make_golden import quiet  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def reference_functions():
  src = open(os.path.join(jax_shim.REFERENCE_SRC, "run_example.py")).read()
  a = src.index("def compute_traj_1d(")
  b = src.index("def solve_HJ(")
  from scipy import interpolate
  ns = {"jnp": np, "np": np, "interpolate": interpolate, "get_f_vals_1d": ref_upd.get_f_vals_1d, "get_f_vals_2d": ref_upd.get_f_vals_2d}
  exec(compile(src[a:b], "run_example.py[18:155]", "exec"), ns)
  return ns["compute_traj_1d"], ns["compute_traj_2d"]


def synth_alp(rng, shape, smooth=True):
  a = rng.standard_normal(shape)
  return a


def main():
  ref1d, ref2d = reference_functions()
  ok = True
  cases = []
  # ---- 1-D: egno 1 linear, egno 2 nearest, with and without noise ----
  for name, egno, nx, nt, ns, epsl, method, seed in (("traj_1d_eg1_linear", 1, 24, 9, 7, 0.0, "linear", 1), ("traj_1d_eg1_noise", 1, 40, 13, 9, 0.05, "linear", 2),
                                                     ("traj_1d_eg2_nearest", 2, 30, 7, 8, 0.02, "nearest", 3)):
    rng = np.random.default_rng(seed)
    P, T = 2.0, 1.0
    x_arr = np.linspace(0.0, P, nx, endpoint=False)
    t_arr = np.linspace(0.0, T, nt)
    alp = rng.standard_normal((2, nt - 1, nx))
    x_init = np.linspace(0.0, P, ns) + (0.3 if epsl else 0.0)      # includes the end point x = period; shifted copy leaves the period
    (fns_ref, _) = quiet(ref_set.set_up_example_fns, egno, 1, 0)
    np.random.seed(100 + seed)
    ta, tx = ref1d(x_init, alp, fns_ref.f_fn, nt, x_arr, t_arr, P, T, epsl, method)
    np.random.seed(100 + seed)
    noise = np.stack([np.random.normal(size=x_init.shape) for _ in range(nt - 1)])
    oa, ox = otr.compute_traj_1d(x_init, alp, orc.set_up_example_fns(egno, 1, 0).f_fn, nt, x_arr, t_arr, P, T, epsl, method, noise=noise)
    err = max(float(np.max(np.abs(oa - ta))), float(np.max(np.abs(ox - tx))))
    ok &= err < 1e-13
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), ndim=1, egno=egno, nx=nx, nt=nt, epsl=epsl, method=method, x_period=P, T=T, x_arr=x_arr,
                        t_arr=t_arr, alp=alp, x_init=x_init, noise=noise, traj_alp=np.asarray(ta), traj_x=np.asarray(tx))
    print("%-24s oracle-vs-reference max abs %.2e" % (name, err), flush=True)
  # ---- 2-D: egno 1 linear (periodic), egno 2 nearest, egno 3 (Neumann-x, centred grids, n_ctrl = 1) ----
  for name, egno, nx, ny, nt, ns, epsl, method, seed in (("traj_2d_eg1_linear", 1, 12, 10, 7, 9, 0.0, "linear", 4), ("traj_2d_eg1_noise", 1, 10, 12, 9, 8, 0.1, "linear", 5),
                                                         ("traj_2d_eg2_nearest", 2, 8, 8, 6, 7, 0.03, "nearest", 6), ("traj_2d_eg3_newton", 3, 10, 12, 8, 6, 0.05, "linear", 7)):
    rng = np.random.default_rng(seed)
    P, T = 2.0, 1.0
    cen = egno == 3
    bc = (1, 0) if egno == 3 else (0, 0)
    n_ctrl = 1 if egno == 3 else 2
    x1 = np.linspace(0.0, P, nx, endpoint=False) - (P / 2 if cen else 0.0)
    x2 = np.linspace(0.0, P, ny, endpoint=False) - (P / 2 if cen else 0.0)
    t_arr = np.linspace(0.0, T, nt)
    alp = rng.standard_normal((4, nt - 1, nx, ny, n_ctrl)) * (1.0 if egno != 3 else 2.0)
    if egno == 3:
      x_init = np.stack([np.full(ns, 0.5), np.linspace(-0.9, 0.9, ns)], axis=-1)
    else:
      x_init = rng.uniform(-0.5, 2.5, (ns, 2))                      # some samples start outside the base period
    (fns_ref, _) = quiet(ref_set.set_up_example_fns, egno, 2, 0)
    np.random.seed(200 + seed)
    ta, tx = ref2d(x_init, alp, fns_ref.f_fn, nt, x1, x2, t_arr, P, P, T, bc, (cen, cen), epsl, method)
    np.random.seed(200 + seed)
    noise = np.stack([np.random.normal(size=x_init.shape) for _ in range(nt - 1)])
    oa, ox = otr.compute_traj_2d(x_init, alp, orc.set_up_example_fns(egno, 2, 0).f_fn, nt, x1, x2, t_arr, P, P, T, bc, (cen, cen), epsl, method,
                                 noise=noise)
    err = max(float(np.max(np.abs(oa - ta))), float(np.max(np.abs(ox - tx))))
    ok &= err < 1e-13
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), ndim=2, egno=egno, nx=nx, ny=ny, nt=nt, epsl=epsl, method=method, x_period=P, y_period=P,
                        T=T, x1_arr=x1, x2_arr=x2, t_arr=t_arr, alp=alp, x_init=x_init, noise=noise, bc=np.array(bc), center=cen,
                        traj_alp=np.asarray(ta), traj_x=np.asarray(tx))
    print("%-24s oracle-vs-reference max abs %.2e" % (name, err), flush=True)
  print("ALL OK" if ok else "MISMATCH")
  return 0 if ok else 1


if __name__ == "__main__":
  sys.exit(main())

"""TEST INFRASTRUCTURE.  Oracle-generated (NumPy restatement, not the reference) fixtures for full-size
BASELINE configs that are too slow to recompute inside the GPU tests.  Usage: python scripts/make_oracle_golden.py cfg1"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pdhg_numpy as orc

CASES = {
  # name: egno, ndim, nx, ny, nt, tsp, epsl, stepsz, N_maxiter, print_freq
  "cfg1": (1, 1, 160, 1, 41, 2, 0.0, 0.1, 1000000, 10000),
  "cfg1_tsp41": (1, 1, 160, 1, 41, 41, 0.0, 0.1, 1000000, 10000),
  "cfg3_blocks3": (1, 2, 256, 256, 4, 2, 0.0, 0.1, 1000000, 10000),   # first 3 blocks of cfg3 need nt=65; see below
}

def main(name):
  egno, ndim, nx, ny, nt, tsp, epsl, stepsz, nmax, pf = CASES[name]
  x_arr, bc, n_ctrl = orc.make_grid(egno, ndim, nx, ny, 2.0, 2.0)
  fns = orc.set_up_example_fns(egno, ndim, 0)
  info, stats = {}, {}
  t0 = time.time()
  res, errs = orc.solve_HJ(ndim, n_ctrl, egno, epsl, fns, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, tsp, stepsz, nmax, pf, 1e-6, bc,
                           info=info, stats=stats)
  dt = time.time() - t0
  mi, phi, rho, alp = res[0]
  print(name, "max_iters", mi, "total iters", sum(info['block_iters']), "time %.1fs" % dt, "inner hist", stats.get('inner_hist'))
  np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_%s.npz" % name), egno=egno, ndim=ndim, nx=nx, ny=ny, nt=nt,
                      tsp=tsp, epsl=epsl, stepsz=stepsz, N_maxiter=nmax, print_freq=pf, max_iters=int(mi), phi=phi, rho=rho, alp=alp,
                      block_iters=np.array(info['block_iters']), stepsz_used=np.array(info['stepsz_used']),
                      errs_nrec=np.array([len(e) for e in errs]), errs_flat=np.concatenate([np.asarray(e).reshape(-1, 2) for e in errs]),
                      oracle_seconds=dt, n_inner=stats['n_inner'])

if __name__ == "__main__":
  main(sys.argv[1])

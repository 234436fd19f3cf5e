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



def cfg3_blocks012():
  """BASELINE configs[2] (256 x 256, nt = 65, tsp = 2, stepsz 0.1), time blocks 0..2, from the NumPy ORACLE (labelled oracle-generated;
  the reference-generated twin is tests/golden/baseline_cfg3_blocks012.npz from oracle/make_golden_baseline.py, ~5 h through the shim)."""
  import numpy as np
  from oracle import pdhg_numpy as orc
  nx = ny = 256
  nblk, nt_full = 3, 65
  nt, T = nblk + 1, nblk / (nt_full - 1.0)
  x_arr, bc, n_ctrl = orc.make_grid(1, 2, nx, ny, 2.0, 2.0)
  info = {}
  res, errs = orc.solve_HJ(2, n_ctrl, 1, 0.0, orc.set_up_example_fns(1, 2, 0), nx, ny, nt, 2.0, 2.0, T, x_arr, 70.0, 2, 0.1, 1000000, 10000, 1e-6, bc,
                           info=info)
  mi, phi, rho, alp = res[0]
  steps = [s for (_, s) in info['stepsz_tried']]
  decr = [steps[i] for i in range(1, len(steps)) if steps[i] != steps[i - 1]]
  s = 4
  alp = np.asarray(alp)
  np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_cfg3_blocks012.npz"), egno=1, ndim=2, nx=nx, ny=ny, nt=nt, T=T, tsp=2, epsl=0.0,
                      stepsz=0.1, N_maxiter=1000000, print_freq=10000, max_iters=int(mi), block_iters=np.array(info['block_iters']),
                      stepsz_used=np.array(info['stepsz_used']), stepsz_decrements=np.array(decr), phi=np.asarray(phi),
                      rho_sub=np.asarray(rho)[:, ::s, ::s], alp_sub=alp[:, :, ::s, ::s, :], sub=s, generated_by="oracle")
  print("cfg3 blocks 0-2 (oracle): iterations", info['block_iters'], "stepsz_used", info['stepsz_used'])


if __name__ == "__main__":
  if sys.argv[1] == "cfg3_blocks012":
    cfg3_blocks012()
  else:
    main(sys.argv[1])

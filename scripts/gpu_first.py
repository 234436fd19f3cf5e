"""Scratch: first GPU run of the 1-D single-CTA kernel against the golden vectors."""
import glob, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pdhg-optimal-control_b200"))
from pdhg_b200 import _lib

def relmax(a, b):
  return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

for f in sorted(glob.glob(os.path.join(ROOT, "tests/golden/*solve_1d*.npz")) + glob.glob(os.path.join(ROOT, "tests/golden/oracle_cfg1.npz"))):
  d = np.load(f)
  if 'failed_block0' in d and bool(d['failed_block0']):
    continue
  egno, nx, nt, tsp = int(d['egno']), int(d['nx']), int(d['nt']), int(d['tsp'])
  K = tsp - 1; nblocks = (nt - 1) // K
  dt, dx = 1.0 / (nt - 1), 2.0 / nx
  x = np.linspace(0.0, 2.0, num=nx, endpoint=False)
  coef = (x - 1.0) ** 2 + 0.1
  C_, pw, Ct = (float(d['C']), float(d['pow']), float(d['Ct'])) if 'C' in d else (1.0, 1.0, 1.0)
  s = _lib.Solver(1, egno, nx, 1, K, 1, 0, dt, dx, 1.0, 70.0, coef, None, C_, pw, Ct, 1e-6, 10, 1, nblocks, 128)
  g = np.sin(2 * np.pi / 2.0 * x)[None, :]
  t0 = time.time()
  phi, rho, alp, logs = s.multi_step_host(g, float(d['epsl']), float(d['stepsz']), int(d['N_maxiter']), int(d['print_freq']))
  el = time.time() - t0
  print(os.path.basename(f), "path", s.path, "status", logs.status, "iters", logs.iters[0].tolist()[:6], "gold", d['block_iters'].tolist()[:6],
        "iters_equal", np.array_equal(logs.iters[0], d['block_iters']), "stepsz", np.array_equal(logs.stepsz_used[0], d['stepsz_used']))
  print("   phi %.2e rho %.2e alp %.2e  time %.3fs  it/s %.0f inner %d" % (relmax(phi[0], d['phi']), relmax(rho[0], d['rho']),
        relmax(alp[0], d['alp']), el, logs.iters.sum() / el, logs.inner_total[0]))
  nrec = logs.nrec[0]
  ok = np.array_equal(nrec, d['errs_nrec'])
  ef = np.concatenate([logs.errlog[0, b, :nrec[b], :2] for b in range(nblocks)])
  print("   nrec equal", ok, "errs rel", relmax(ef, d['errs_flat']) if ok else None)
  s.close()

"""Scratch: cooperative kernel + operator entry points against the golden vectors."""
import glob, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pdhg-optimal-control_b200"))
import contextlib, io
import pdhg_b200
from pdhg_b200 import update_fns_in_pdhg as upd, run_example as rx, set_fns

def relmax(a, b):
  a = np.asarray(a); b = np.asarray(b)
  return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

def quiet(fn, *a, **k):
  buf = io.StringIO()
  with contextlib.redirect_stdout(buf):
    return fn(*a, **k)

only = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob(os.path.join(ROOT, "tests/golden/op_*.npz"))):
  if only and only not in f: continue
  d = np.load(f)
  egno, ndim, nx, ny, nt, K = [int(d[k]) for k in ("egno", "ndim", "nx", "ny", "nt", "K")]
  if egno == 3: continue
  n_ctrl, bc, cen = rx.problem_setup(egno, ndim)
  x_arr = rx.make_x_arr(ndim, nx, ny, 2.0, 2.0, cen)
  fns = quiet(set_fns.set_up_example_fns, egno, ndim, 0)
  dt = float(d['dt']); dsp = (2.0 / nx,) if ndim == 1 else (2.0 / nx, 2.0 / ny)
  alp = tuple(d['alp'][j] for j in range(2 * ndim))
  fn = upd.update_primal_1d if ndim == 1 else upd.update_primal_2d
  pn = fn(d['phi'], d['rho'], 70.0, alp, float(d['tau']), dt, dsp, fns, None, float(d['epsl']), x_arr, None, bc, C=float(d['C']), pow=float(d['pow']), Ct=float(d['Ct']))
  r1, a1, e1 = upd.update_dual_oneiter(d['phi_bar'], d['rho'], 70.0, alp, float(d['sigma']), dt, dsp, float(d['epsl']), x_arr, None, bc, fns, ndim)
  rN, aN = upd.update_dual_alternative(d['phi_bar'], d['rho'], 70.0, alp, float(d['sigma']), dt, dsp, float(d['epsl']), fns, x_arr, None, ndim, bc, eps=float(d['eps']))
  print("%-26s primal %.1e | sweep1 rho %.1e alp %.1e err %.1e | dual rho %.1e alp %.1e" % (
    os.path.basename(f), relmax(pn, d['phi_next']), relmax(r1, d['rho_sweep1']), relmax(np.stack(a1), d['alp_sweep1']),
    abs(e1 - float(d['err_sweep1'])) / abs(float(d['err_sweep1'])), relmax(rN, d['rho_dual']), relmax(np.stack(aN), d['alp_dual'])), flush=True)

for f in sorted(glob.glob(os.path.join(ROOT, "tests/golden/solve_*.npz"))):
  if only and only not in f: continue
  d = np.load(f)
  egno, ndim, nx, ny, nt, tsp = [int(d[k]) for k in ("egno", "ndim", "nx", "ny", "nt", "tsp")]
  if egno == 3 or bool(d['failed_block0']): continue
  n_ctrl, bc, cen = rx.problem_setup(egno, ndim)
  x_arr = rx.make_x_arr(ndim, nx, ny, 2.0, 2.0, cen)
  fns = quiet(set_fns.set_up_example_fns, egno, ndim, 0)
  for path in ([1, 2] if ndim == 1 else [2]):
    upd.clear_handles()
    os.environ["PDHG_FORCE_PATH"] = str(path)
    info = {}
    t0 = time.time()
    res, errs = quiet(rx.solve_HJ, ndim, n_ctrl, egno, float(d['epsl']), fns, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, tsp, float(d['stepsz']),
                      int(d['N_maxiter']), int(d['print_freq']), 1e-6, bc, C=float(d['C']), pow=float(d['pow']), Ct=float(d['Ct']), info=info)
    el = time.time() - t0
    mi, phi, rho, alp = res[0]
    ef = np.concatenate([e.reshape(-1, 2) for e in errs])
    print("%-32s path %d iters_eq %s stepsz_eq %s phi %.1e rho %.1e alp %.1e errs %s  %.2fs %.0f it/s" % (
      os.path.basename(f), info['path'], info['block_iters'] == d['block_iters'].tolist(), info['stepsz_used'] == d['stepsz_used'].tolist(),
      relmax(phi, d['phi']), relmax(rho, d['rho']), relmax(alp, d['alp']),
      ("%.1e" % relmax(ef, d['errs_flat'])) if ef.shape == d['errs_flat'].shape else "shape!", el, sum(info['block_iters']) / el), flush=True)

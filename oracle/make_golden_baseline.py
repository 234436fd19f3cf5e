"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the two BASELINE behaviours recorded in SURVEY.md Appendix C, produced by the
UNMODIFIED reference sources (through oracle/jax_shim.py, like oracle/make_golden.py) and cross-checked against the NumPy
restatement:

  baseline_cfg2_stepsz01_failed   configs[1]: egno=1 ndim=1 epsl=0.1 nx=640 nt=161 at stepsz_param=0.1 — every step size of the
                                  fallback chain 0.1 -> 0.01 NaNs in block 0 and the solve ends in "algorithm failed"
                                  (utils_pdhg_solver.py:180-187)
  baseline_cfg3_blocks012         configs[2]: egno=1 ndim=2 epsl=0 nx=ny=256 nt=65 tsp=2 stepsz_param=0.1, time blocks 0..2 —
                                  block 1 NaNs at 0.1, 0.09, 0.08 and converges at 0.07000000000000002; per-block iteration counts

Runs in the BUILD container only (needs /root/reference); ~25 minutes of CPU for the second case.

    python oracle/make_golden_baseline.py [cfg2|cfg3]
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_golden as mg  # noqa: E402  (installs the shim, imports the reference modules)

orc, ref_sol, ref_set, ref_pre = mg.orc, mg.ref_sol, mg.ref_set, mg.ref_pre
GOLD = mg.GOLD


def run_reference(egno, ndim, nx, ny, nt, T, tsp, epsl, stepsz, nmax, pf):
  x_arr, bc, n_ctrl, dt, dspatial, nspatial, period = mg.grids(egno, ndim, nx, ny, nt, T=T)
  (fns_ref, _) = mg.quiet(ref_set.set_up_example_fns, egno, ndim, 0)
  g_ref = ref_set.set_up_J(egno, ndim, period)(x_arr)
  fv_ref = ref_pre.compute_Dxx_fft_fv(ndim, nspatial, dspatial, bc)
  prim_ref, dual_ref = mg.ref_lambdas(ndim, bc, 1.0, 1.0, 1.0)
  buf = io.StringIO()
  res, errs, crashed = None, [], False
  t0 = time.time()
  try:
    with contextlib.redirect_stdout(buf), np.errstate(all='ignore'):
      res, errs = ref_sol.PDHG_multi_step(prim_ref, dual_ref, fns_ref, g_ref, x_arr, ndim, nt, nspatial, dt, dspatial, 70.0,
                                          time_step_per_PDHG=tsp, epsl=epsl, stepsz_param=stepsz, n_ctrl=n_ctrl, fv=fv_ref,
                                          N_maxiter=nmax, print_freq=pf, eps=1e-6)
  except (ValueError, UnboundLocalError):
    crashed = True     # the reference crashes after a block-0 failure (utils_pdhg_solver.py:206/215); its log lines are kept
  log = buf.getvalue()
  return res, errs, log, crashed, time.time() - t0, (x_arr, bc, n_ctrl)


def parse_log(log):
  steps = [float(l.rsplit(' ', 1)[1]) for l in log.splitlines() if 'decrease step size to' in l]
  nan_iters = [int(l.split('iter ')[1]) for l in log.splitlines() if l.startswith('Nan error at iter')]
  conv = [int(l.split('iter ')[1]) + 1 for l in log.splitlines() if l.startswith('PDHG converges at iter')]
  return steps, nan_iters, conv, ('algorithm failed' in log)


def cfg2():
  egno, ndim, nx, ny, nt, tsp, epsl, stepsz, nmax, pf = 1, 1, 640, 1, 161, 2, 0.1, 0.1, 1000000, 10000
  res, errs, log, crashed, t_ref, (x_arr, bc, n_ctrl) = run_reference(egno, ndim, nx, ny, nt, 1.0, tsp, epsl, stepsz, nmax, pf)
  steps, nan_iters, conv, failed = parse_log(log)
  info = {}
  orc.solve_HJ(ndim, n_ctrl, egno, epsl, orc.set_up_example_fns(egno, ndim, 0), nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, tsp, stepsz,
               nmax, pf, 1e-6, bc, info=info)
  orc_steps = [s for (_, s) in info['stepsz_tried']]
  orc_decr = [orc_steps[i] for i in range(1, len(orc_steps)) if orc_steps[i] != orc_steps[i - 1]]
  assert steps == orc_decr, (steps, orc_decr)
  assert failed and info['sol_nan'] and not conv
  np.savez_compressed(os.path.join(GOLD, "baseline_cfg2_stepsz01_failed.npz"), egno=egno, ndim=ndim, nx=nx, ny=ny, nt=nt, tsp=tsp,
                      epsl=epsl, stepsz=stepsz, N_maxiter=nmax, print_freq=pf, failed_block0=True, sol_nan=True,
                      stepsz_decrements=np.array(steps), nan_iters=np.array(nan_iters), reference_crashed=crashed)
  print("cfg2 @0.1: decrements %s, NaN at iterations %s, algorithm failed=%s (reference %.1fs)" % (steps, nan_iters, failed, t_ref), flush=True)


def cfg3():
  egno, ndim, nx, ny, nt_full, tsp, epsl, stepsz, nmax, pf = 1, 2, 256, 256, 65, 2, 0.0, 0.1, 1000000, 10000
  nblk = 3
  nt, T = nblk + 1, nblk / (nt_full - 1.0)        # same dt as the full nt=65, T=1 run: blocks 0..2 of it
  res, errs, log, crashed, t_ref, (x_arr, bc, n_ctrl) = run_reference(egno, ndim, nx, ny, nt, T, tsp, epsl, stepsz, nmax, pf)
  assert not crashed
  steps, nan_iters, conv, failed = parse_log(log)
  print("cfg3 reference: block iterations %s, decrements %s, NaN at %s (%.0fs)" % (conv, steps, nan_iters, t_ref), flush=True)
  t0 = time.time()
  info = {}
  res_o, errs_o = orc.solve_HJ(ndim, n_ctrl, egno, epsl, orc.set_up_example_fns(egno, ndim, 0), nx, ny, nt, 2.0, 2.0, T, x_arr, 70.0,
                               tsp, stepsz, nmax, pf, 1e-6, bc, info=info)
  mi, phi, rho, alp = res[0]
  mi_o, phi_o, rho_o, alp_o = res_o[0]
  worst = max(mg.relmax(phi_o, phi), mg.relmax(rho_o, rho), mg.relmax(alp_o, alp))
  print("cfg3 oracle: block iterations %s, stepsz_used %s, oracle-vs-reference rel-Linf %.2e (%.0fs)"
        % (info['block_iters'], info['stepsz_used'], worst, time.time() - t0), flush=True)
  assert list(info['block_iters']) == conv and int(mi) == int(mi_o)
  orc_steps = [s for (_, s) in info['stepsz_tried']]
  orc_decr = [orc_steps[i] for i in range(1, len(orc_steps)) if orc_steps[i] != orc_steps[i - 1]]
  assert steps == orc_decr, (steps, orc_decr)
  nrec = np.array([len(e) for e in errs])
  errs_flat = np.concatenate([np.asarray(e).reshape(-1, 2) for e in errs], axis=0)
  alp = np.asarray(alp)
  s = 4     # rho / alp are stored on every 4th grid line (fixture size); phi in full
  np.savez_compressed(os.path.join(GOLD, "baseline_cfg3_blocks012.npz"), egno=egno, ndim=ndim, nx=nx, ny=ny, nt=nt, T=T, tsp=tsp,
                      epsl=epsl, stepsz=stepsz, N_maxiter=nmax, print_freq=pf, max_iters=int(mi), block_iters=np.array(conv),
                      stepsz_used=np.array(info['stepsz_used']), stepsz_decrements=np.array(steps), nan_iters=np.array(nan_iters),
                      phi=np.asarray(phi), rho_sub=np.asarray(rho)[:, ::s, ::s], alp_sub=alp[:, :, ::s, ::s, :], sub=s,
                      rho_sum=float(np.sum(rho)), alp_abs_sum=float(np.sum(np.abs(alp))), errs_nrec=nrec, errs_flat=errs_flat,
                      oracle_vs_reference=worst)


if __name__ == "__main__":
  which = sys.argv[1] if len(sys.argv) > 1 else "all"
  if which in ("cfg2", "all"):
    cfg2()
  if which in ("cfg3", "all"):
    cfg3()

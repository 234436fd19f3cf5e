"""TEST INFRASTRUCTURE ONLY.  Generates `tests/golden/*.npz` and validates `oracle/pdhg_numpy.py`.

Runs in the BUILD container only (needs `/root/reference`): the unmodified reference sources are
imported through `oracle/jax_shim.py` (NumPy-backed stand-in for the jax API; jaxlib is not installed
and cannot be), executed on small deterministic inputs, and their outputs are
  (1) compared with the NumPy restatement (must agree to rounding), and
  (2) written to `tests/golden/` so the GPU box (which has no `/root/reference`) can test against them.

    python oracle/make_golden.py            # regenerate + validate
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

from oracle import jax_shim  # noqa: E402

jax_shim.install()

import update_fns_in_pdhg as ref_upd            # noqa: E402  (reference, via shim)
import set_fns as ref_set                       # noqa: E402
from utils import utils_precond as ref_pre      # noqa: E402
from utils import utils_pdhg_solver as ref_sol  # noqa: E402

from oracle import pdhg_numpy as orc            # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def quiet(fn, *a, **k):
  buf = io.StringIO()
  with contextlib.redirect_stdout(buf), np.errstate(all='ignore'):
    out = fn(*a, **k)
  return out, buf.getvalue()


def relmax(a, b):
  a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
  if a.size == 0:
    return 0.0
  if not np.array_equal(np.isnan(a), np.isnan(b)):
    return np.inf
  m = ~np.isnan(a)
  den = max(np.max(np.abs(b[m])), 1e-300) if m.any() else 1.0
  return float(np.max(np.abs(a[m] - b[m])) / den) if m.any() else 0.0


def grids(egno, ndim, nx, ny, nt, T=1.0, x_period=2.0, y_period=2.0):
  x_arr, bc, n_ctrl = orc.make_grid(egno, ndim, nx, ny, x_period, y_period)
  dt = T / (nt - 1)
  if ndim == 1:
    dspatial, nspatial, period = (x_period / nx,), (nx,), (x_period,)
  else:
    dspatial, nspatial, period = (x_period / nx, y_period / ny), (nx, ny), (x_period, y_period)
  return x_arr, bc, n_ctrl, dt, dspatial, nspatial, period


def random_state(rng, K, nspatial, ndim, n_ctrl, egno):
  """A mid-trajectory-like random state: rho >= 0 with some zeros, signed alp, smooth-ish phi."""
  sh = (K,) + tuple(nspatial)
  phi = rng.standard_normal((K + 1,) + tuple(nspatial))
  rho = np.maximum(rng.standard_normal(sh) * 30 + 40, 0.0)
  alp = []
  for j in range(2 * ndim):
    a = rng.standard_normal(sh + (n_ctrl,))
    if egno != 3 and ndim == 2:           # structurally-zero components stay zero in the reference
      a[..., 1 if j < 2 else 0] = 0.0
    if egno == 3 and j >= 2:
      a[...] = 0.0
    alp.append(a)
  return phi, rho, tuple(alp)


def ref_lambdas(ndim, bc, C, pow, Ct):
  """The two closures of run_example.py:193-203."""
  if ndim == 1:
    prim = lambda phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr: \
      ref_upd.update_primal_1d(phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc,
                               C=C, pow=pow, Ct=Ct)
  else:
    prim = lambda phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr: \
      ref_upd.update_primal_2d(phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc,
                               C=C, pow=pow, Ct=Ct)
  dual = lambda phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr, ndim, eps: \
    ref_upd.update_dual_alternative(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr,
                                    ndim, bc, eps=eps)
  return prim, dual


OP_CASES = [
  # name, egno, ndim, nx, ny, nt(grid for dt), K, epsl, C, pow, Ct
  ("op_1d_eg1_K1",      1, 1, 24, 1, 13, 1, 0.0, 1.0, 1.0, 1.0),
  ("op_1d_eg1_K1_visc", 1, 1, 20, 1, 11, 1, 0.1, 1.0, 1.0, 1.0),
  ("op_1d_eg1_K4_visc", 1, 1, 24, 1, 13, 4, 0.05, 1.0, 1.0, 1.0),
  ("op_1d_eg1_K3_pow2", 1, 1, 30, 1, 7, 3, 0.02, 0.5, 2.0, 0.7),
  ("op_1d_eg1_K2_Ct0",  1, 1, 16, 1, 9, 2, 0.0, 1.0, 1.0, 0.0),
  ("op_1d_eg2_K2",      2, 1, 24, 1, 13, 2, 0.03, 1.0, 1.0, 1.0),
  ("op_2d_eg1_K1",      1, 2, 12, 10, 9, 1, 0.0, 1.0, 1.0, 1.0),
  ("op_2d_eg1_K3_visc", 1, 2, 8, 12, 7, 3, 0.1, 1.0, 1.0, 1.0),
  ("op_2d_eg2_K2",      2, 2, 10, 8, 5, 2, 0.02, 1.0, 1.0, 1.0),
  ("op_2d_eg3_K1",      3, 2, 10, 12, 9, 1, 0.05, 1.0, 1.0, 1.0),
]


def run_op_case(name, egno, ndim, nx, ny, nt, K, epsl, C, pw, Ct, seed):
  rng = np.random.default_rng(seed)
  x_arr, bc, n_ctrl, dt, dspatial, nspatial, _ = grids(egno, ndim, nx, ny, nt)
  phi, rho, alp = random_state(rng, K, nspatial, ndim, n_ctrl, egno)
  t_arr = np.linspace(0, dt * K, K + 1)[1:]
  t_arr = t_arr[:, None] if ndim == 1 else t_arr[:, None, None]
  c_on_rho, tau, sigma, eps = 70.0, 0.1 / 1.5, 0.1 * 1.5, 1e-6
  (fns_ref, _) = quiet(ref_set.set_up_example_fns, egno, ndim, 0)
  fns_orc = orc.set_up_example_fns(egno, ndim, 0)
  fv_ref = ref_pre.compute_Dxx_fft_fv(ndim, nspatial, dspatial, bc)
  fv_orc = orc.compute_Dxx_fft_fv(ndim, nspatial, dspatial, bc)
  prim_ref, dual_ref = ref_lambdas(ndim, bc, C, pw, Ct)
  up_orc = orc.update_primal_1d if ndim == 1 else orc.update_primal_2d
  with np.errstate(all='ignore'):
    phi_next_ref = prim_ref(phi, rho, c_on_rho, alp, tau, dt, dspatial, fns_ref, fv_ref, epsl, x_arr, t_arr)
    phi_next_orc = up_orc(phi, rho, c_on_rho, alp, tau, dt, dspatial, fns_orc, fv_orc, epsl, x_arr, t_arr, bc, C=C, pow=pw, Ct=Ct)
    res_ref = (ref_upd.compute_cont_residual_1d if ndim == 1 else ref_upd.compute_cont_residual_2d)(
      rho, alp, dt, dspatial, fns_ref, c_on_rho, epsl, x_arr, t_arr, bc)
    res_orc = orc.compute_cont_residual(rho, alp, dt, dspatial, fns_orc, c_on_rho, epsl, x_arr, t_arr, bc)
    phi_bar = 2 * phi_next_ref - phi
    r1_ref, a1_ref, e1_ref = ref_upd.update_dual_oneiter(phi_bar, rho, c_on_rho, alp, sigma, dt, dspatial, epsl, x_arr, t_arr, bc, fns_ref, ndim)
    r1_orc, a1_orc, e1_orc = orc.update_dual_oneiter(phi_bar, rho, c_on_rho, alp, sigma, dt, dspatial, epsl, x_arr, t_arr, bc, fns_orc, ndim)
    rN_ref, aN_ref = dual_ref(phi_bar, rho, c_on_rho, alp, sigma, dt, dspatial, epsl, fns_ref, x_arr, t_arr, ndim, eps)
    st = {}
    rN_orc, aN_orc = orc.update_dual_alternative(phi_bar, rho, c_on_rho, alp, sigma, dt, dspatial, epsl, fns_orc, x_arr, t_arr,
                                                 ndim, bc, eps=eps, stats=st)
  worst = max(relmax(fv_orc.real, fv_ref.real), relmax(res_orc, res_ref), relmax(phi_next_orc, phi_next_ref),
              relmax(r1_orc, r1_ref), relmax(np.stack(a1_orc), np.stack(a1_ref)), relmax(e1_orc, e1_ref),
              relmax(rN_orc, rN_ref), relmax(np.stack(aN_orc), np.stack(aN_ref)))
  np.savez_compressed(
    os.path.join(GOLD, name + ".npz"),
    egno=egno, ndim=ndim, nx=nx, ny=ny, nt=nt, K=K, epsl=epsl, C=C, pow=pw, Ct=Ct, c_on_rho=c_on_rho, tau=tau, sigma=sigma,
    eps=eps, dt=dt, phi=phi, rho=rho, alp=np.stack(alp), cont_residual=res_ref, phi_next=phi_next_ref, phi_bar=phi_bar,
    rho_sweep1=r1_ref, alp_sweep1=np.stack(a1_ref), err_sweep1=e1_ref, rho_dual=rN_ref, alp_dual=np.stack(aN_ref),
    n_inner=st['n_inner'], fv_real=np.real(fv_ref))
  return worst


SOLVE_CASES = [
  # name, egno, ndim, nx, ny, nt, tsp, epsl, stepsz, N_maxiter, print_freq, (C,pow,Ct)
  ("solve_1d_eg1_nx40_nt11",        1, 1, 40, 1, 11, 2, 0.0, 0.1, 1000000, 1000, (1.0, 1.0, 1.0)),
  ("solve_1d_eg1_nx20_nt6_tsp6",    1, 1, 20, 1, 6, 6, 0.0, 0.1, 1000000, 500, (1.0, 1.0, 1.0)),
  ("solve_1d_eg1_nx40_nt11_visc",   1, 1, 40, 1, 11, 2, 0.1, 0.1, 1000000, 10000, (1.0, 1.0, 1.0)),   # NaN -> step-size fallback
  ("solve_1d_eg1_nx64_nt21_fail",   1, 1, 64, 1, 21, 2, 0.3, 0.1, 3000, 10000, (1.0, 1.0, 1.0)),      # fallback (and possibly failure)
  ("solve_1d_eg2_nx30_nt7",         2, 1, 30, 1, 7, 2, 0.0, 0.1, 1000000, 2000, (1.0, 1.0, 1.0)),
  ("solve_1d_eg1_nx24_nt9_tsp3_pow",1, 1, 24, 1, 9, 3, 0.01, 0.1, 4000, 1000, (0.5, 2.0, 0.5)),
  ("solve_2d_eg1_12x12_nt4",        1, 2, 12, 12, 4, 2, 0.0, 0.1, 1000000, 1000, (1.0, 1.0, 1.0)),
  ("solve_2d_eg1_10x8_nt5_tsp3",    1, 2, 10, 8, 5, 3, 0.05, 0.05, 600, 200, (1.0, 1.0, 1.0)),
  ("solve_2d_eg2_8x8_nt3",          2, 2, 8, 8, 3, 2, 0.0, 0.1, 800, 300, (1.0, 1.0, 1.0)),
  ("solve_2d_eg3_10x12_nt3",        3, 2, 10, 12, 3, 2, 0.1, 0.05, 500, 100, (1.0, 1.0, 1.0)),
  ("solve_1d_eg1_nx32_nt5_failed",  1, 1, 32, 1, 5, 2, 5.0, 0.1, 400, 10000, (1.0, 1.0, 1.0)),   # every step size NaNs: "algorithm failed"
]


def run_solve_case(name, egno, ndim, nx, ny, nt, tsp, epsl, stepsz, nmax, pf, pre):
  C, pw, Ct = pre
  x_arr, bc, n_ctrl, dt, dspatial, nspatial, period = grids(egno, ndim, nx, ny, nt)
  (fns_ref, _) = quiet(ref_set.set_up_example_fns, egno, ndim, 0)
  fns_orc = orc.set_up_example_fns(egno, ndim, 0)
  g_ref = ref_set.set_up_J(egno, ndim, period)(x_arr)
  fv_ref = ref_pre.compute_Dxx_fft_fv(ndim, nspatial, dspatial, bc)
  prim_ref, dual_ref = ref_lambdas(ndim, bc, C, pw, Ct)
  t0 = time.time()
  failed = False
  try:
    ((res_ref, errs_ref), log) = quiet(
      ref_sol.PDHG_multi_step, prim_ref, dual_ref, fns_ref, g_ref, x_arr, ndim, nt, nspatial, dt, dspatial, 70.0,
      time_step_per_PDHG=tsp, epsl=epsl, stepsz_param=stepsz, n_ctrl=n_ctrl, fv=fv_ref, N_maxiter=nmax, print_freq=pf, eps=1e-6)
  except (ValueError, UnboundLocalError):
    # reference crashes (UnboundLocalError on `pdhg_iters` at :206, else concatenate([]) when block 0 fails (utils_pdhg_solver.py:215); its log up to the crash is lost with
    # the redirected stdout, so the fallback announcements are re-captured line by line below
    failed, res_ref, errs_ref, log = True, None, [], ""
    buf = io.StringIO()
    try:
      with contextlib.redirect_stdout(buf), np.errstate(all='ignore'):
        ref_sol.PDHG_multi_step(prim_ref, dual_ref, fns_ref, g_ref, x_arr, ndim, nt, nspatial, dt, dspatial, 70.0, time_step_per_PDHG=tsp,
                                epsl=epsl, stepsz_param=stepsz, n_ctrl=n_ctrl, fv=fv_ref, N_maxiter=nmax, print_freq=pf, eps=1e-6)
    except (ValueError, UnboundLocalError):
      log = buf.getvalue()
  t_ref = time.time() - t0
  info = {}
  res_orc, errs_orc = orc.solve_HJ(ndim, n_ctrl, egno, epsl, fns_orc, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, tsp, stepsz, nmax, pf,
                                   1e-6, bc, C=C, pow=pw, Ct=Ct, info=info)
  # step sizes announced by the reference's own log lines
  ref_steps = [float(l.rsplit(' ', 1)[1]) for l in log.splitlines() if 'decrease step size to' in l]
  ref_failed = ('algorithm failed' in log) or failed
  orc_steps = [s for (_, s) in info['stepsz_tried']]
  orc_decr = [orc_steps[i] for i in range(1, len(orc_steps)) if orc_steps[i] != orc_steps[i - 1]]
  assert ref_steps == orc_decr, (name, ref_steps, orc_decr)
  assert ref_failed == info['sol_nan'], name
  if failed:
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), egno=egno, ndim=ndim, nx=nx, ny=ny, nt=nt, tsp=tsp, epsl=epsl,
                        stepsz=stepsz, N_maxiter=nmax, print_freq=pf, C=C, pow=pw, Ct=Ct, failed_block0=True,
                        stepsz_decrements=np.array(ref_steps), sol_nan=True)
    return 0.0, t_ref, 0
  mi_ref, phi_ref, rho_ref, alp_ref = res_ref[0]
  mi_orc, phi_orc, rho_orc, alp_orc = res_orc[0]
  assert int(mi_ref) == int(mi_orc), (name, mi_ref, mi_orc)
  assert len(errs_ref) == len(errs_orc)
  block_iters_ref = [int(l.split('iter ')[1]) + 1 for l in log.splitlines() if l.startswith('PDHG converges at iter')]
  worst = max(relmax(phi_orc, phi_ref), relmax(rho_orc, rho_ref), relmax(alp_orc, alp_ref))
  for a, b in zip(errs_orc, errs_ref):
    assert np.shape(a) == np.shape(b), (name, np.shape(a), np.shape(b))
    worst = max(worst, relmax(a, b) if np.size(a) else 0.0)
  nrec = np.array([len(e) for e in errs_ref])
  errs_flat = np.concatenate([np.asarray(e).reshape(-1, 2) for e in errs_ref], axis=0) if len(errs_ref) else np.zeros((0, 2))
  np.savez_compressed(
    os.path.join(GOLD, name + ".npz"), egno=egno, ndim=ndim, nx=nx, ny=ny, nt=nt, tsp=tsp, epsl=epsl, stepsz=stepsz,
    N_maxiter=nmax, print_freq=pf, C=C, pow=pw, Ct=Ct, failed_block0=False, max_iters=int(mi_ref), phi=phi_ref, rho=rho_ref,
    alp=alp_ref, errs_nrec=nrec, errs_flat=errs_flat, block_iters=np.array(info['block_iters']),
    stepsz_used=np.array(info['stepsz_used']), stepsz_decrements=np.array(ref_steps), sol_nan=bool(ref_failed))
  return worst, t_ref, int(mi_ref)


def main():
  os.makedirs(GOLD, exist_ok=True)
  ok = True
  only = sys.argv[1] if len(sys.argv) > 1 else None     # optional substring filter: regenerate selected cases only
  global OP_CASES, SOLVE_CASES
  if only:
    OP_CASES = [c for c in OP_CASES if only in c[0]]
    SOLVE_CASES = [c for c in SOLVE_CASES if only in c[0]]
  for seed, case in enumerate(OP_CASES):
    w = run_op_case(*case, seed=100 + seed)
    print("%-32s oracle-vs-reference rel-Linf %.2e" % (case[0], w), flush=True)
    ok &= w < 1e-11
  for case in SOLVE_CASES:
    w, t, mi = run_solve_case(*case)
    print("%-32s oracle-vs-reference rel-Linf %.2e  max_iters %d  (reference via shim: %.1fs)" % (case[0], w, mi, t), flush=True)
    ok &= w < 1e-9
  print("ALL OK" if ok else "MISMATCH")
  return 0 if ok else 1


if __name__ == "__main__":
  sys.exit(main())

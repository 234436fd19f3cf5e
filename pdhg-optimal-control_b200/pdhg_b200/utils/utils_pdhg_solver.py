"""PDHG loop and time-block marching (mirror of jaxsrc/utils/utils_pdhg_solver.py).

`PDHG_solver_oneiter` and `PDHG_multi_step` keep the reference's signatures and return values.  When the two
injected callables are the native ones (`NativeUpdatePrimal` / `NativeUpdateDual`, which is what `solve_HJ`
builds) the whole loop — iterations, inner dual sweeps, error norms, convergence / NaN exits, time-block
hand-off and the step-size fallback — runs inside CUDA kernels (`pdhg_solve_block` / `pdhg_multi_step`) and
the host only formats the log lines afterwards.  With any other callables the reference's host loop is
reproduced verbatim around them (tensors stay on the GPU; norms are torch reductions).
"""
import numpy as np

from .. import _dev, _lib
from ..solver import save
from ..update_fns_in_pdhg import NativeUpdateDual, NativeUpdatePrimal, get_solver
from . import utils


def _native(fn_update_primal, fn_update_dual):
  return isinstance(fn_update_primal, NativeUpdatePrimal) and isinstance(fn_update_dual, NativeUpdateDual)


def _max_rec(N_maxiter, print_freq):
  if print_freq is None or print_freq <= 0:
    return 2
  return int(min(N_maxiter // print_freq + 3, 1 << 20))


def PDHG_solver_oneiter(fn_update_primal, fn_update_dual, fns_dict, phi0, rho0, alp0, x_arr, t_arr,
                        ndim, dt, dspatial, c_on_rho, epsl=0.0, stepsz_param=0.9, fv=None,
                        N_maxiter=1000000, print_freq=1000, eps=1e-6, tfboard=False, tfrecord_ind=0, snapshots=True):
  """One time block of PDHG (utils_pdhg_solver.py:9-94).
  Returns (results_all, error_all): results_all = [(i, phi_prev, rho_prev, alp_next) at every i % print_freq == 0 ...,
  (pdhg_iters, phi_next, rho_next, alp_next)], error_all [n_records, 2].  `snapshots=False` (extension) skips the
  periodic state snapshots (their error rows are still recorded) and runs the block in a single kernel launch."""
  if not _native(fn_update_primal, fn_update_dual):
    return _host_loop(fn_update_primal, fn_update_dual, fns_dict, phi0, rho0, alp0, x_arr, t_arr, ndim, dt, dspatial, c_on_rho,
                      epsl, stepsz_param, fv, N_maxiter, print_freq, eps)
  t = _dev.require_cuda()
  K = rho0.shape[0]
  P = fn_update_primal
  pf = int(print_freq) if print_freq else 0
  s = get_solver(fns_dict, rho0.shape[1:], K, P.bc, dt, dspatial, c_on_rho, x_arr, C=P.C, pow=P.pow, Ct=P.Ct, eps=eps,
                 rho_alp_iters=fn_update_dual.rho_alp_iters, max_rec=_max_rec(N_maxiter, pf) if not snapshots else 4)
  phi = _dev.to_dev(phi0)
  rho = _dev.to_dev(rho0)
  alp = t.stack([_dev.to_dev(a) for a in alp0], dim=0).contiguous()
  out = lambda d, like: _dev.like_input(d, like)
  unstack = lambda a: tuple(out(a[j], alp0[0]) for j in range(a.shape[0]))
  results_all, error_all = [], []
  i = 0
  reason = _lib.END_MAXITER
  while True:
    single = snapshots and pf > 0 and i % pf == 0
    pause = (i + 1) if single else ((min((i // pf + 1) * pf, N_maxiter) if (snapshots and pf > 0) else N_maxiter))
    phi_n, rho_n, alp_n = t.empty_like(phi), t.empty_like(rho), t.empty_like(alp)
    logs = s.solve_block_dev(phi.data_ptr(), rho.data_ptr(), alp.data_ptr(), float(epsl), float(stepsz_param), int(N_maxiter),
                             i, pause if pause < N_maxiter else 0, pf, phi_n.data_ptr(), rho_n.data_ptr(), alp_n.data_ptr(),
                             _dev.stream_ptr())
    reason = int(logs.end_reason[0, 0])
    nrec = int(logs.nrec[0, 0])
    rows = logs.errlog[0, 0, :nrec]
    ended = reason != _lib.END_PAUSED
    n_periodic = nrec - 1 if ended else nrec
    for r in range(n_periodic):
      it_r = i if single else (i + ((-i) % pf) + r * pf)
      if snapshots:
        results_all.append((it_r, out(phi, phi0), out(rho, rho0), unstack(alp_n)))
      error_all.append(rows[r, :2].copy())
      print('iteration {}, primal error {:.2E}, dual error {:.2E}, min rho {:.2f}, max rho {:.2f}'.format(
        it_r, rows[r, 0], rows[r, 1], rows[r, 2], rows[r, 3]), flush=True)
    i = int(logs.iters[0, 0])
    phi, rho, alp = phi_n, rho_n, alp_n
    if ended:
      last = rows[nrec - 1]
      if reason == _lib.END_CONVERGED:
        print('PDHG converges at iter {}'.format(i - 1), flush=True)
      elif reason == _lib.END_NAN:
        print("Nan error at iter {}".format(i - 1))
      # the reference formats error[2] here, which JAX clamps to error[1] (utils_pdhg_solver.py:90)
      print('iteration {}, primal error with prev step {:.2E}, dual error with prev step {:.2E}, eqt error {:.2E}'.format(
        i - 1, last[0], last[1], last[1]), flush=True)
      results_all.append((i, out(phi, phi0), out(rho, rho0), unstack(alp)))
      error_all.append(last[:2].copy())
      break
  return results_all, np.array(error_all)


def _host_loop(fn_update_primal, fn_update_dual, fns_dict, phi0, rho0, alp0, x_arr, t_arr, ndim, dt, dspatial, c_on_rho,
               epsl, stepsz_param, fv, N_maxiter, print_freq, eps):
  """Reference loop around arbitrary injected callables (utils_pdhg_solver.py:40-94); state lives on the GPU."""
  t = _dev.require_cuda()
  phi_prev, rho_prev = _dev.to_dev(phi0), _dev.to_dev(rho0)
  alp_prev = tuple(_dev.to_dev(a) for a in alp0)
  scale = 1.5
  tau_phi, tau_rho = stepsz_param / scale, stepsz_param * scale
  norm = lambda a: float(t.linalg.vector_norm(a))
  error_all, results_all = [], []
  error = np.array([np.nan, np.nan])
  i = -1
  for i in range(N_maxiter):
    phi_next = _dev.to_dev(fn_update_primal(phi_prev, rho_prev, c_on_rho, alp_prev, tau_phi, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr))
    phi_bar = 2 * phi_next - phi_prev
    rho_next, alp_next = fn_update_dual(phi_bar, rho_prev, c_on_rho, alp_prev, tau_rho, dt, dspatial, epsl, fns_dict, x_arr, t_arr,
                                        ndim, eps=eps)
    rho_next = _dev.to_dev(rho_next)
    alp_next = tuple(_dev.to_dev(a) for a in alp_next)
    with np.errstate(all='ignore'):
      err1 = np.float64(norm(phi_next - phi_prev)) / np.float64(norm(phi_prev))
      err2 = np.float64(norm(rho_next - rho_prev)) / np.float64(norm(rho_prev))
      for alp_p, alp_n in zip(alp_prev, alp_next):
        norm_alp, norm_err = norm(alp_p), norm(alp_p - alp_n)
        if norm_alp < 1e-6 and norm_err > 1e-6:
          err2 += norm_err
        elif norm_alp >= 1e-6:
          err2 += norm_err / norm_alp
    error = np.array([err1, err2])
    if error[0] < eps and error[1] < eps:
      print('PDHG converges at iter {}'.format(i), flush=True)
      break
    if bool(t.isnan(phi_next).any()) or bool(t.isnan(rho_next).any()):
      print("Nan error at iter {}".format(i))
      break
    if print_freq > 0 and i % print_freq == 0:
      results_all.append((i, _dev.like_input(phi_prev, phi0), _dev.like_input(rho_prev, rho0),
                          tuple(_dev.like_input(a, alp0[0]) for a in alp_next)))
      error_all.append(error)
      print('iteration {}, primal error {:.2E}, dual error {:.2E}, min rho {:.2f}, max rho {:.2f}'.format(
        i, error[0], error[1], float(rho_next.min()), float(rho_next.max())), flush=True)
    phi_prev, rho_prev, alp_prev = phi_next, rho_next, alp_next
  print('iteration {}, primal error with prev step {:.2E}, dual error with prev step {:.2E}, eqt error {:.2E}'.format(
    i, error[0], error[1], error[1]), flush=True)
  results_all.append((i + 1, _dev.like_input(phi_next, phi0), _dev.like_input(rho_next, rho0),
                      tuple(_dev.like_input(a, alp0[0]) for a in alp_next)))
  error_all.append(error)
  return results_all, np.array(error_all)


def _print_block_log(i, nt_PDHG, logs, b, pf):
  """Replays the reference's per-block log lines from the device-side records."""
  print('=================== nt_PDHG = {}, i = {} ==================='.format(nt_PDHG, i), flush=True)
  nrec, rows = int(logs.nrec[b, i]), logs.errlog[b, i]
  iters, reason = int(logs.iters[b, i]), int(logs.end_reason[b, i])
  for r in range(nrec - 1):
    print('iteration {}, primal error {:.2E}, dual error {:.2E}, min rho {:.2f}, max rho {:.2f}'.format(
      r * pf, rows[r, 0], rows[r, 1], rows[r, 2], rows[r, 3]), flush=True)
  if reason == _lib.END_CONVERGED:
    print('PDHG converges at iter {}'.format(iters - 1), flush=True)
  last = rows[nrec - 1]
  print('iteration {}, primal error with prev step {:.2E}, dual error with prev step {:.2E}, eqt error {:.2E}'.format(
    iters - 1, last[0], last[1], last[1]), flush=True)


def PDHG_multi_step(fn_update_primal, fn_update_dual, fns_dict, g, x_arr,
                    ndim, nt, nspatial, dt, dspatial, c_on_rho, time_step_per_PDHG=2,
                    epsl=0.0, stepsz_param=0.9, n_ctrl=None, fv=None,
                    N_maxiter=1000000, print_freq=1000, eps=1e-6, tfboard=False,
                    save_middle_dir=None, save_middle_prefix=None,
                    load_middle_dir=None, load_middle_prefix=None, info=None):
  """Time-block marching (utils_pdhg_solver.py:97-225).
  Returns ([(max_iters, phi[nt,...], rho[nt-1,...], alp[2 ndim, nt-1, ..., n_ctrl])], errs_all).
  `info` (extension): dict receiving per-block iteration counts, step sizes and kernel statistics."""
  if n_ctrl is None:
    n_ctrl = ndim
  tsp = time_step_per_PDHG
  assert (nt - 1) % (tsp - 1) == 0
  nt_PDHG = (nt - 1) // (tsp - 1)
  K = tsp - 1
  if not _native(fn_update_primal, fn_update_dual):
    raise NotImplementedError("PDHG_multi_step runs the fused CUDA march and needs the native update callables "
                              "(update_fns_in_pdhg.NativeUpdatePrimal / NativeUpdateDual); for custom callables drive "
                              "PDHG_solver_oneiter block by block")
  P = fn_update_primal
  pf = int(print_freq) if print_freq else 0
  print('shape of phi0: ', (tsp,) + tuple(nspatial), flush=True)
  print('shape of rho0: ', (K,) + tuple(nspatial), flush=True)
  print('shape of alp0: ', (2 * ndim, K) + tuple(nspatial) + (n_ctrl,), flush=True)
  s = get_solver(fns_dict, nspatial, K, P.bc, dt, dspatial, c_on_rho, x_arr, C=P.C, pow=P.pow, Ct=P.Ct, eps=eps,
                 rho_alp_iters=fn_update_dual.rho_alp_iters, batch=1, nblocks=nt_PDHG, max_rec=_max_rec(N_maxiter, pf))
  utils.timer.tic("time estimate")
  on_dev = _dev.is_tensor(g)
  if on_dev:
    t = _dev.require_cuda()
    g_d = _dev.to_dev(g)
    phi_d = t.empty((1, nt) + tuple(nspatial), dtype=t.float64, device=g_d.device)
    rho_d = t.empty((1, nt - 1) + tuple(nspatial), dtype=t.float64, device=g_d.device)
    alp_d = t.empty((1, 2 * ndim, nt - 1) + tuple(nspatial) + (n_ctrl,), dtype=t.float64, device=g_d.device)
    logs = s.multi_step_dev(g_d.data_ptr(), float(epsl), float(stepsz_param), int(N_maxiter), pf, phi_d.data_ptr(),
                            rho_d.data_ptr(), alp_d.data_ptr(), _dev.stream_ptr())
    phi, rho, alp = phi_d, rho_d, alp_d
  else:
    phi, rho, alp, logs = s.multi_step_host(np.asarray(g, dtype=np.float64).reshape((1,) + tuple(nspatial)), float(epsl),
                                            float(stepsz_param), int(N_maxiter), pf)
  done = int(logs.blocks_done[0])
  sol_nan = int(logs.status[0]) == _lib.INST_SOL_NAN
  if int(logs.status[0]) == _lib.INST_LOG_OVERFLOW:
    print('warning: error log overflow, some periodic records were dropped', flush=True)
  # replay of the log lines (the march itself is one kernel launch)
  step = float(stepsz_param)
  errs_all = []
  for i in range(done):
    used = float(logs.stepsz_used[0, i])
    while step != used:   # fallback announcements of utils_pdhg_solver.py:181-183
      step -= float(stepsz_param) / 10
      print('pdhg does not conv at t_ind = {}, decrease step size to {}'.format(i, step), flush=True)
    _print_block_log(i, nt_PDHG, logs, 0, pf)
    errs_all.append(logs.errlog[0, i, :int(logs.nrec[0, i]), :2].copy())
  if sol_nan:
    while step > float(logs.stepsz_final[0]):
      step -= float(stepsz_param) / 10
      print('pdhg does not conv at t_ind = {}, decrease step size to {}'.format(done, step), flush=True)
    print('pdhg does not conv at t_ind = {}, algorithm failed'.format(done), flush=True)
  utils.timer.estimate_time("time estimate", max(done, 1) / nt_PDHG, int(logs.iters[0, :done].sum()) if done else None)
  max_iters = int(logs.iters[0, :done].max()) if done else 0
  if info is not None:
    info.update(block_iters=logs.iters[0, :done].tolist(), stepsz_used=logs.stepsz_used[0, :done].tolist(), sol_nan=sol_nan,
                stepsz_final=float(logs.stepsz_final[0]), n_inner=int(logs.inner_total[0]), blocks_done=done,
                path=s.path, launches=s.launch_count)
  if done == 0:
    # the reference raises in jnp.concatenate([]) here (utils_pdhg_solver.py:215); report the failure instead
    print('pdhg does not conv, please decrease stepsize to be less than {}'.format(float(logs.stepsz_final[0])), flush=True)
    return [(0, None, None, None)], errs_all
  rows = done * K + (1 if done == nt_PDHG else 0)
  phi_out, rho_out, alp_out = phi[0, :rows], rho[0, :done * K], alp[0, :, :done * K]
  results_out = [(max_iters, phi_out, rho_out, alp_out)]
  if save_middle_dir is not None and save_middle_prefix is not None:
    Kk = K
    save(save_middle_dir, save_middle_prefix,
         [max_iters, [phi_out[i * Kk:(i + 1) * Kk + (1 if i == nt_PDHG - 1 else 0)] for i in range(done)],
          [rho_out[i * Kk:(i + 1) * Kk] for i in range(done)], [alp_out[:, i * Kk:(i + 1) * Kk] for i in range(done)], errs_all])
  print('\n\n')
  print('===========================================')
  if sol_nan:
    print('pdhg does not conv, please decrease stepsize to be less than {}'.format(float(logs.stepsz_final[0])), flush=True)
  else:
    print('pdhg conv. Max err is {:.2E}. Max iters is {}'.format(max(float(np.max(e)) for e in errs_all), max_iters), flush=True)
  return results_out, errs_all

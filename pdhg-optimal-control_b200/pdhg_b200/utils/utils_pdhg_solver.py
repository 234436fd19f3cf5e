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
  if _dev.is_tensor(alp0[0]):
    alp = t.stack([_dev.to_dev(a) for a in alp0], dim=0).contiguous()
  else:   # host arrays go straight into their slice of the stacked device array (no second pass over 4 x the state)
    alp = t.empty((len(alp0),) + tuple(alp0[0].shape), dtype=t.float64, device=phi.device)
    for j, a_ in enumerate(alp0):
      alp[j].copy_(t.from_numpy(np.ascontiguousarray(np.asarray(a_, dtype=np.float64))), non_blocking=True)
  out = lambda d, like: _dev.like_input(d, like)

  def unstack(a):
    if _dev.is_tensor(alp0[0]):
      return tuple(a[j] for j in range(a.shape[0]))
    h = _dev.to_host([a])[0]             # one pinned copy of the stacked array; the tuple entries are views of it
    return tuple(h[j] for j in range(h.shape[0]))
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


def _middle_lists(phi, rho, alp, done, K, nt_PDHG):
  """The reference's middle-file lists (utils_pdhg_solver.py:193-199): per block phi_curr[:-1] (the last block: all of it), rho, alp."""
  return ([phi[i * K:(i + 1) * K + (1 if i == nt_PDHG - 1 else 0)] for i in range(done)], [rho[i * K:(i + 1) * K] for i in range(done)],
          [alp[:, i * K:(i + 1) * K] for i in range(done)])


def _march_blockwise(s, g, nt, nt_PDHG, K, ndim, nspatial, n_ctrl, epsl, stepsz_param, N_maxiter, pf, save_dir, save_prefix, load_dir, load_prefix):
  """One pdhg_multi_step_range launch per time block.  After every block the middle file
      [max_iters, phi_all, rho_all, alp_all, errs_all, resume]
  is rewritten: entries 0..4 are the reference's (solver.py:28-33), `resume` = {"phi0": the next block's warm-started phi0,
  "stepsz_param": the step size after the fallbacks so far, "blocks_done"} is what a restart needs and the reference's file lacks
  (it drops phi_curr[-1] of the last saved block and the current step size, which is why its load path cannot work)."""
  from ..solver import load_middle_solution
  t = _dev.require_cuda()
  dev = "cuda:%d" % _dev.current_device()
  nsp = tuple(nspatial)
  A = 2 * ndim
  phi = t.zeros((1, nt) + nsp, dtype=t.float64, device=dev)
  rho = t.zeros((1, nt - 1) + nsp, dtype=t.float64, device=dev)
  alp = t.zeros((1, A, nt - 1) + nsp + (n_ctrl,), dtype=t.float64, device=dev)
  g_d = _dev.to_dev(np.asarray(g, dtype=np.float64).reshape((1,) + nsp) if not _dev.is_tensor(g) else g.reshape((1,) + nsp))
  logs = _lib.LogBuffers(1, nt_PDHG, s.max_rec)
  begin, cur = 0, float(stepsz_param)
  if load_dir is not None:
    mid = load_middle_solution(load_dir, load_prefix)
    begin = len(mid[1])
    assert begin == len(mid[2]) == len(mid[3]) == len(mid[4])
    if begin > 0:
      if len(mid) < 6 or not isinstance(mid[5], dict) or "phi0" not in mid[5]:
        raise ValueError("middle file {}/{} has no resume record (written by the reference, whose load path is unwired): cannot restart "
                         "from it".format(load_dir, load_prefix))
      rs = mid[5]
      assert int(rs["blocks_done"]) == begin
      cur = float(rs["stepsz_param"])
      for i in range(begin):      # re-install the blocks already solved
        pb = np.asarray(mid[1][i]); nr = pb.shape[0]
        phi[0, i * K:i * K + nr] = _dev.to_dev(pb)
        rho[0, i * K:(i + 1) * K] = _dev.to_dev(np.asarray(mid[2][i]))
        alp[0, :, i * K:(i + 1) * K] = _dev.to_dev(np.asarray(mid[3][i]))
        e = np.asarray(mid[4][i], dtype=np.float64).reshape(-1, 2)
        logs.nrec[0, i] = len(e); logs.errlog[0, i, :len(e), :2] = e
        logs.iters[0, i] = int(rs["block_iters"][i]); logs.stepsz_used[0, i] = float(rs["stepsz_used"][i]); logs.end_reason[0, i] = int(rs["end_reason"][i])
      rho0 = rho[0, (begin - 1) * K:begin * K].contiguous()
      alp0 = alp[0, :, (begin - 1) * K:begin * K].contiguous()
      phi0 = _dev.to_dev(np.asarray(rs["phi0"]))
      s.set_march_state(phi0.data_ptr(), rho0.data_ptr(), alp0.data_ptr(), _dev.stream_ptr())
      logs.blocks_done[0] = begin; logs.stepsz_final[0] = cur
  kept = {k: getattr(logs, k).copy() for k in ("iters", "stepsz_used", "nrec", "errlog", "end_reason")}
  inner = 0
  for i in range(begin, nt_PDHG):
    lg = s.multi_step_range_dev(g_d.data_ptr(), float(epsl), float(stepsz_param), cur, int(N_maxiter), pf, i, i + 1, phi.data_ptr(),
                                rho.data_ptr(), alp.data_ptr(), stream=_dev.stream_ptr())
    inner += int(lg.inner_total[0])
    ok = int(lg.blocks_done[0]) == i + 1
    for k in kept:
      kept[k][0, i] = getattr(lg, k)[0, i]
    cur = float(lg.stepsz_final[0])
    logs.status[0] = lg.status[0]; logs.stepsz_final[0] = cur
    if not ok:
      break
    logs.blocks_done[0] = i + 1
    if save_dir is not None:
      done = i + 1
      phi0 = t.empty((1, K + 1) + nsp, dtype=t.float64, device=dev)
      r0, a0 = t.empty((1, K) + nsp, dtype=t.float64, device=dev), t.empty((1, A, K) + nsp + (n_ctrl,), dtype=t.float64, device=dev)
      s.get_march_state(phi0.data_ptr(), r0.data_ptr(), a0.data_ptr(), _dev.stream_ptr())
      ph, rh, ah = phi[0].cpu().numpy(), rho[0].cpu().numpy(), alp[0].cpu().numpy()
      pl, rl, al = _middle_lists(ph, rh, ah, done, K, nt_PDHG)
      errs = [kept["errlog"][0, j, :int(kept["nrec"][0, j]), :2].copy() for j in range(done)]
      resume = {"phi0": phi0[0].cpu().numpy(), "stepsz_param": cur, "blocks_done": done, "block_iters": kept["iters"][0, :done].tolist(),
                "stepsz_used": kept["stepsz_used"][0, :done].tolist(), "end_reason": kept["end_reason"][0, :done].tolist()}
      save(save_dir, save_prefix, [int(kept["iters"][0, :done].max()), pl, rl, al, errs, resume])
  for k in kept:
    getattr(logs, k)[...] = kept[k]
  logs.inner_total[0] = inner
  return phi, rho, alp, logs


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
  save_mid = save_middle_dir is not None and save_middle_prefix is not None
  load_mid = load_middle_dir is not None and load_middle_prefix is not None
  if save_mid or load_mid:
    # block-wise march (one launch per time block): the middle file is rewritten after EVERY block as in the reference
    # (utils_pdhg_solver.py:211-212), and a march can restart from it (a working version of :139-154)
    phi, rho, alp, logs = _march_blockwise(s, g, nt, nt_PDHG, K, ndim, nspatial, n_ctrl, epsl, stepsz_param, N_maxiter, pf,
                                           save_middle_dir if save_mid else None, save_middle_prefix, load_middle_dir if load_mid else None,
                                           load_middle_prefix)
    if not on_dev:
      phi, rho, alp = phi.cpu().numpy(), rho.cpu().numpy(), alp.cpu().numpy()
  elif on_dev:
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
  if done < nt_PDHG and not sol_nan:
    raise RuntimeError("PDHG_multi_step: the march stopped after {} of {} time blocks without a NaN failure (status {})".format(
      done, nt_PDHG, int(logs.status[0])))
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
                path=s.path, launches=s.launch_count, kernel_ms=s.last_kernel_ms)
  if done == 0:
    # the reference raises in jnp.concatenate([]) here (utils_pdhg_solver.py:215); report the failure instead
    print('pdhg does not conv, please decrease stepsize to be less than {}'.format(float(logs.stepsz_final[0])), flush=True)
    return [(0, None, None, None)], errs_all
  rows = done * K + (1 if done == nt_PDHG else 0)
  phi_out, rho_out, alp_out = phi[0, :rows], rho[0, :done * K], alp[0, :, :done * K]
  results_out = [(max_iters, phi_out, rho_out, alp_out)]
  print('\n\n')
  print('===========================================')
  if sol_nan:
    print('pdhg does not conv, please decrease stepsize to be less than {}'.format(float(logs.stepsz_final[0])), flush=True)
  else:
    print('pdhg conv. Max err is {:.2E}. Max iters is {}'.format(max(float(np.max(e)) for e in errs_all), max_iters), flush=True)
  return results_out, errs_all

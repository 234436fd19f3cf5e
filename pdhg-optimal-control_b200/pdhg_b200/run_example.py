"""Driver with the reference's flag surface (mirror of jaxsrc/run_example.py:157-303,402-442).

    python -m pdhg_b200.run_example --ndim 1 --epsl 0 --egno 1 --nx 160 --nt 41 --stepsz_param 0.1

Kept: every solver flag, grid construction, `solve_HJ`, pickle save/load.  Not kept (out of the hot path,
SURVEY.md section 8): plotting, TensorBoard, closed-loop trajectory simulation.
"""
import os
import sys
from datetime import datetime

import numpy as np

from .set_fns import set_up_example_fns, set_up_J
from .solver import load_solution, save
from .update_fns_in_pdhg import NativeUpdateDual, NativeUpdatePrimal, get_solver
from .utils.utils_precond import compute_Dxx_fft_fv
from .utils.utils_pdhg_solver import PDHG_multi_step, _max_rec
from . import _lib


def problem_setup(egno, ndim):
  """bc / n_ctrl / grid centring by example (run_example.py:228-240)."""
  if egno == 3:
    assert ndim == 2
    return 1, (1, 0), True
  return ndim, (0 if ndim == 1 else (0, 0)), False


def make_x_arr(ndim, nx, ny, x_period, y_period, centered=False):
  """x_arr [1,nx,1] or [1,nx,ny,2] (run_example.py:273-287)."""
  if ndim == 1:
    x_arr = np.linspace(0.0, x_period, num=nx, endpoint=False)[None, :, None]
    return x_arr - x_period / 2 if centered else x_arr
  x1 = np.linspace(0.0, x_period, num=nx, endpoint=False)
  x2 = np.linspace(0.0, y_period, num=ny, endpoint=False)
  if centered:
    x1, x2 = x1 - x_period / 2, x2 - y_period / 2
  xm, ym = np.meshgrid(x1, x2, indexing='ij')
  return np.stack([xm, ym], axis=-1)[None, ...]


def _grid(ndim, nx, ny, nt, x_period, y_period, T):
  dt = T / (nt - 1)
  dx = x_period / nx
  dy = y_period / ny
  if ndim == 1:
    return dt, (x_period,), (dx,), (nx,)
  return dt, (x_period, y_period), (dx, dy), (nx, ny)


def solve_HJ(ndim, n_ctrl, egno, epsl, fns_dict, nx, ny, nt, x_period, y_period, T, x_arr,
             c_on_rho, time_step_per_PDHG, stepsz_param, N_maxiter, print_freq, eps, bc, save_dir=None,
             save_middle_dir=None, save_middle_prefix=None, C=1.0, pow=1.0, Ct=1.0, info=None, load_middle_dir=None,
             load_middle_prefix=None):
  """run_example.py:157-210.  `C`, `pow`, `Ct` replace the reference's reads of FLAGS inside the closures."""
  dt, period_spatial, dspatial, nspatial = _grid(ndim, nx, ny, nt, x_period, y_period, T)
  print('period_spatial: ', period_spatial)
  print('dspatial: ', dspatial)
  print('nspatial: ', nspatial)
  g = set_up_J(egno, ndim, period_spatial)(x_arr)
  print('shape of g: ', g.shape)
  fv = compute_Dxx_fft_fv(ndim, nspatial, dspatial, bc) if (bc == 0 or tuple(np.atleast_1d(bc)) == (0, 0)) else None
  fn_update_primal = NativeUpdatePrimal(ndim, bc, C=C, pow=pow, Ct=Ct)
  fn_update_dual = NativeUpdateDual(bc)
  return PDHG_multi_step(fn_update_primal, fn_update_dual, fns_dict, g, x_arr, ndim, nt, nspatial, dt, dspatial, c_on_rho,
                         time_step_per_PDHG=time_step_per_PDHG, epsl=epsl, stepsz_param=stepsz_param, fv=fv, n_ctrl=n_ctrl,
                         N_maxiter=N_maxiter, print_freq=print_freq, eps=eps, save_middle_dir=save_middle_dir,
                         save_middle_prefix=save_middle_prefix, load_middle_dir=load_middle_dir, load_middle_prefix=load_middle_prefix,
                         info=info)


def solve_HJ_batch(ndim, n_ctrl, egno, epsl, fns_dict, nx, ny, nt, x_period, y_period, T, x_arr, g,
                   c_on_rho, time_step_per_PDHG, stepsz_param, N_maxiter, print_freq, eps, bc,
                   C=1.0, pow=1.0, Ct=1.0, device=0, path=0, info=None):
  """Extension for sweeps: B independent instances (initial data g[B,...], epsl[B], stepsz_param[B]) solved by ONE
  launch, one CTA (or CTA group) per instance; no collective.  Returns (phi[B,nt,..], rho, alp, logs)."""
  dt, _, dspatial, nspatial = _grid(ndim, nx, ny, nt, x_period, y_period, T)
  g = np.asarray(g, dtype=np.float64)
  B = g.shape[0]
  K = time_step_per_PDHG - 1
  assert (nt - 1) % K == 0
  s = get_solver(fns_dict, nspatial, K, bc, dt, dspatial, c_on_rho, x_arr, C=C, pow=pow, Ct=Ct, eps=eps, batch=B,
                 nblocks=(nt - 1) // K, max_rec=_max_rec(N_maxiter, print_freq), device=device, path=path)
  out = s.multi_step_host(g, epsl, stepsz_param, N_maxiter, print_freq)
  if info is not None:
    info.update(kernel_ms=s.last_kernel_ms, path=s.path, launches=s.launch_count)
  return out


def _traj_noise(epsl, shape, seed, noise):
  """Standard normal increments [nt-1, n_sample(, 2)] in numpy.random's order (the reference draws them step by step from the
  global stream, run_example.py:48,151); `seed` makes a run reproducible, `noise` injects the draws themselves."""
  if epsl <= 0:
    return None
  if noise is not None:
    return np.ascontiguousarray(np.asarray(noise, dtype=np.float64).reshape(shape))
  rs = np.random.RandomState(seed) if seed is not None else np.random
  return np.stack([rs.normal(size=shape[1:]) for _ in range(shape[0])])


def compute_traj_1d(x_init, alp, fns_dict, nt, x_arr, t_arr, x_period, T, epsl=0.0, interp_method='linear', seed=None, noise=None):
  """run_example.py:18-51 on the GPU (pdhg_compute_traj): alp [2, nt-1, nx], x_init [n_sample] ->
  (traj_alp [nt-1, n_sample, 1], traj_x [nt, n_sample]).  `fns_dict` (instead of the reference's bare f_fn) names the dynamics."""
  from . import _dev
  t = _dev.require_cuda()
  x_init = np.asarray(_np_(x_init), dtype=np.float64).ravel()
  xa = np.asarray(_np_(x_arr), dtype=np.float64).ravel()
  ns, nx = x_init.size, xa.size
  alp_d = _dev.to_dev(alp)
  if interp_method == 'nearest':
    assert np.all(np.diff(xa) > 0), "nearest interpolation expects an ascending grid"      # (:38-39 compares with x_arr itself)
  else:
    xa = xa % x_period                                         # numpy.interp(period=): nodes reduced mod period and sorted
    order = np.argsort(xa, kind="stable")
    if not np.array_equal(order, np.arange(nx)):
      xa = xa[order]
      alp_d = alp_d[:, :, t.from_numpy(order).to(alp_d.device)].contiguous()
  xa_dev = _dev.to_dev(xa)
  nz = _traj_noise(epsl, (nt - 1, ns), seed, noise)
  nz_d = _dev.to_dev(nz) if nz is not None else None
  tx = t.empty((nt, ns), dtype=t.float64, device=alp_d.device)
  ta = t.empty((nt - 1, ns, 1), dtype=t.float64, device=alp_d.device)
  x0_d, t_d = _dev.to_dev(x_init), _dev.to_dev(np.asarray(_np_(t_arr), dtype=np.float64).ravel())
  _lib.compute_traj_dev(1, fns_dict.egno, 1, nx, 1, nt, ns, 0, 0, interp_method == 'nearest', x_period, 1.0, epsl, alp_d.data_ptr(), xa_dev.data_ptr(),
                        None, t_d.data_ptr(), nz_d.data_ptr() if nz_d is not None else None, x0_d.data_ptr(), tx.data_ptr(), ta.data_ptr(),
                        _dev.stream_ptr())
  return (ta, tx) if _dev.is_tensor(alp) else tuple(_dev.to_host([ta, tx]))


def compute_traj_2d(x_init, alp, fns_dict, nt, x1_arr, x2_arr, t_arr, x_period, y_period, T, bc, center, epsl=0.0, interp_method='linear',
                    seed=None, noise=None):
  """run_example.py:113-155 on the GPU: alp [4, nt-1, nx, ny, n_ctrl], x_init [n_sample, 2] ->
  (traj_alp [nt-1, n_sample, n_ctrl], traj_x [nt, n_sample, 2]).  bc (0, 0) periodic, (1, 0) edge-clamped in x (egno 3)."""
  from . import _dev
  t = _dev.require_cuda()
  x_init = np.ascontiguousarray(np.asarray(_np_(x_init), dtype=np.float64).reshape(-1, 2))
  ns = x_init.shape[0]
  alp_d = _dev.to_dev(_np_(alp) if not _dev.is_tensor(alp) else alp)
  nx, ny, n_ctrl = int(alp_d.shape[2]), int(alp_d.shape[3]), int(alp_d.shape[4])
  nz = _traj_noise(epsl, (nt - 1, ns, 2), seed, noise)
  nz_d = _dev.to_dev(nz) if nz is not None else None
  tx = t.empty((nt, ns, 2), dtype=t.float64, device=alp_d.device)
  ta = t.empty((nt - 1, ns, n_ctrl), dtype=t.float64, device=alp_d.device)
  x1_d, x2_d = _dev.to_dev(np.asarray(_np_(x1_arr), dtype=np.float64).ravel()), _dev.to_dev(np.asarray(_np_(x2_arr), dtype=np.float64).ravel())
  x0_d, t_d = _dev.to_dev(x_init), _dev.to_dev(np.asarray(_np_(t_arr), dtype=np.float64).ravel())
  _lib.compute_traj_dev(2, fns_dict.egno, n_ctrl, nx, ny, nt, ns, bc[0], bc[1], interp_method == 'nearest', x_period, y_period, epsl, alp_d.data_ptr(),
                        x1_d.data_ptr(), x2_d.data_ptr(), t_d.data_ptr(), nz_d.data_ptr() if nz_d is not None else None, x0_d.data_ptr(),
                        tx.data_ptr(), ta.data_ptr(), _dev.stream_ptr())
  return (ta, tx) if _dev.is_tensor(alp) else tuple(_dev.to_host([ta, tx]))


def _np_(x):
  return x.detach().cpu().numpy() if hasattr(x, "detach") else x


def main(argv):
  from absl import flags
  FLAGS = flags.FLAGS
  for key, value in FLAGS.__flags.items():
    print(value.name, ": ", value._value, flush=True)
  nt, nx, ny, ndim, egno = FLAGS.nt, FLAGS.nx, FLAGS.ny, FLAGS.ndim, FLAGS.egno
  n_ctrl, bc, centered = problem_setup(egno, ndim)
  if ndim == 1:
    filename_prefix = 'nt{}_nx{}'.format(nt, nx)
  elif ndim == 2:
    filename_prefix = 'nt{}_nx{}_ny{}'.format(nt, nx, ny)
  else:
    raise NotImplementedError
  if FLAGS.load or FLAGS.load_middle:
    assert FLAGS.load_timestamp != ''
    time_stamp = FLAGS.load_timestamp
  else:
    time_stamp = datetime.now().strftime("%Y%m%d-%H%M%S")
  save_dir = './check_points/{}'.format(time_stamp) + '/eg{}_{}d'.format(egno, ndim)
  fns_dict = set_up_example_fns(egno, ndim, FLAGS.numerical_L_ind)
  x_arr = make_x_arr(ndim, nx, ny, FLAGS.x_period, FLAGS.y_period, centered)
  if FLAGS.load:
    results, errs_all = load_solution(save_dir, filename_prefix)
  else:
    smd, smp = (save_dir, filename_prefix) if FLAGS.save_middle else (None, None)
    # --load_middle --load_timestamp T restarts the march from the middle file of run T (written by --save_middle) at the first
    # unsolved time block; the reference defines the flag but never wires it (run_example.py:249-254)
    # (same file name as the final pickle, as in the reference :293-295: an interrupted run leaves the middle list there)
    lmd, lmp = (save_dir, filename_prefix) if FLAGS.load_middle else (None, None)
    results, errs_all = solve_HJ(ndim, n_ctrl, egno, FLAGS.epsl, fns_dict, nx, ny, nt, FLAGS.x_period, FLAGS.y_period, FLAGS.T,
                                 x_arr, FLAGS.c_on_rho, FLAGS.time_step_per_PDHG, FLAGS.stepsz_param, FLAGS.N_maxiter,
                                 FLAGS.print_freq, FLAGS.eps, bc, save_middle_dir=smd, save_middle_prefix=smp,
                                 C=FLAGS.C, pow=FLAGS.pow, Ct=FLAGS.Ct, load_middle_dir=lmd, load_middle_prefix=lmp)
    if FLAGS.save:
      save(save_dir, filename_prefix, (results, errs_all))
  if FLAGS.plot_traj_num_1d > 0 and results[0][3] is not None:
    # closed-loop trajectories (run_example.py:342-393); the figures themselves are out of scope, the arrays are saved
    alp = np.asarray(_np_(results[0][3]))
    alp_rev = alp[:, ::-1]                                       # :356 time direction of the control
    t_arr = np.linspace(0.0, FLAGS.T, num=nt)
    n1 = FLAGS.plot_traj_num_1d
    method = 'nearest' if egno == 2 else 'linear'
    if egno == 3:
      ys = np.linspace(-FLAGS.y_period / 2 + 0.1, FLAGS.y_period / 2 - 0.1, num=n1)[:, None]
      if FLAGS.epsl > 0:
        ys = 0 * ys
      xs = np.pad(ys, ((0, 0), (1, 0)), mode='constant', constant_values=0.5)
      traj_alp, traj_x = compute_traj_2d(xs, alp_rev, fns_dict, nt, x_arr[0, :, 0, 0], x_arr[0, 0, :, 1], t_arr, FLAGS.x_period, FLAGS.y_period,
                                         FLAGS.T, bc, (centered, centered), FLAGS.epsl)
    elif ndim == 1:
      traj_alp, traj_x = compute_traj_1d(np.linspace(0, FLAGS.x_period, num=n1), alp_rev[..., 0], fns_dict, nt, x_arr[0, :, 0], t_arr,
                                         FLAGS.x_period, FLAGS.T, FLAGS.epsl, method)
    else:
      xm, ym = np.meshgrid(np.linspace(0, FLAGS.x_period, num=n1), np.linspace(0, FLAGS.y_period, num=n1), indexing='ij')
      traj_alp, traj_x = compute_traj_2d(np.stack([xm.flatten(), ym.flatten()], axis=-1), alp_rev, fns_dict, nt, x_arr[0, :, 0, 0], x_arr[0, 0, :, 1],
                                         t_arr, FLAGS.x_period, FLAGS.y_period, FLAGS.T, bc, (centered, centered), FLAGS.epsl, method)
    save(save_dir, filename_prefix + '_traj', (traj_alp, traj_x))
  if FLAGS.plot or FLAGS.tfboard:
    print('plotting / tensorboard are outside this package (SURVEY.md section 8)')
  print('phi: ', results[0][1])
  print('end')


def define_flags():
  from absl import flags
  flags.DEFINE_integer('egno', 1, 'index of example, corresponding to the three examples in the paper')
  flags.DEFINE_integer('ndim', 1, 'spatial dimension')
  flags.DEFINE_float('epsl', 0.0, 'diffusion coefficient')
  flags.DEFINE_float('x_period', 2.0, 'period of x')
  flags.DEFINE_float('y_period', 2.0, 'period of y')
  flags.DEFINE_integer('nt', 11, 'size of t grids')
  flags.DEFINE_integer('nx', 20, 'size of x grids')
  flags.DEFINE_integer('ny', 20, 'size of y grids')
  flags.DEFINE_float('stepsz_param', 0.1, 'step sizes in PDHG')
  flags.DEFINE_boolean('save', True, 'if save the final results to a pickle file')
  flags.DEFINE_boolean('save_middle', False, 'if save middle results')
  flags.DEFINE_boolean('load', False, 'if load the final results from the pickle file')
  flags.DEFINE_boolean('load_middle', False, 'if load middle results')
  flags.DEFINE_string('load_timestamp', '', 'the timestamp of the folder to load from')
  flags.DEFINE_boolean('tfboard', False, 'if use tfboard for plotting')
  flags.DEFINE_boolean('plot', False, 'true if plot the figures of phi and alp, and the trajectories')
  flags.DEFINE_integer('plot_traj_num_1d', 0, 'number of trajectories to plot')
  flags.DEFINE_float('T', 1.0, 'time horizon')
  flags.DEFINE_float('c_on_rho', 70.0, 'the constant c in the objective function')
  flags.DEFINE_integer('time_step_per_PDHG', 2, 'number of time discretization per PDHG iteration')
  flags.DEFINE_integer('N_maxiter', 1000000, 'maximum number of iterations')
  flags.DEFINE_integer('print_freq', 10000, 'print frequency')
  flags.DEFINE_float('eps', 1e-6, 'the error threshold')
  flags.DEFINE_float('C', 1.0, 'constant in preconditioning')
  flags.DEFINE_float('pow', 1.0, 'power in preconditioning')
  flags.DEFINE_float('Ct', 1.0, 'constant in preconditioning')
  flags.DEFINE_integer('numerical_L_ind', 0, 'index of numerical L')


if __name__ == '__main__':
  from absl import app
  define_flags()
  app.run(main)

"""Problem definitions (mirror of jaxsrc/set_fns.py).

In the reference `set_up_example_fns` returns three JAX closures that are traced into the jitted update
operators.  Closures cannot cross a C ABI, so here the namedtuple additionally carries the problem id
(`egno`, `ndim`, `n_ctrl`) that selects the compiled-in CUDA implementation of the same formulas
(csrc/pdhg_device.cuh: prox_alp / f_plus / f_minus / lagr), and `coef_tables()` evaluates the coefficient
a(x) = (x-1)^2 + 0.1 on the grid for the kernels.  The three callables are kept as plain NumPy functions
for host-side post-processing (e.g. trajectory simulation); the solver never calls them.
"""
from collections import namedtuple

import numpy as np

Functions = namedtuple('Functions', ['f_fn', 'numerical_L_fn', 'alp_update_fn', 'egno', 'ndim', 'n_ctrl'])


def set_up_J(egno, ndim, period_spatial):
  """Initial condition g (set_fns.py:10-24)."""
  if egno != 3:
    if ndim == 1:
      alpha = 2 * np.pi / period_spatial[0]
    elif ndim == 2:
      alpha = np.array([2 * np.pi / period_spatial[0], 2 * np.pi / period_spatial[1]])
    else:
      raise ValueError("ndim {} not implemented".format(ndim))
    J = lambda x: np.sum(np.sin(alpha * np.asarray(x)), axis=-1)
  else:
    x_period, y_period = period_spatial
    J = lambda x: np.sin(2 * np.pi / y_period * np.asarray(x)[..., 1]) * np.exp(-np.asarray(x)[..., 0] ** 2 / 2)
  return J


def coef_tables(egno, ndim, x_arr):
  """Per-direction coefficient tables passed to pdhg_create: a(x), a(y) for egno 1,2 (set_fns.py:117-118,145);
  the velocity grid x itself for egno 3 (f_y = x, set_fns.py:98)."""
  x_arr = np.asarray(x_arr, dtype=np.float64)
  a = lambda v: (v - 1.0) ** 2 + 0.1
  if ndim == 1:
    return a(x_arr[0, :, 0]), None
  xs, ys = x_arr[0, :, 0, 0], x_arr[0, 0, :, 1]
  if egno == 3:
    return xs.copy(), ys.copy()
  return a(xs), a(ys)


def set_up_example_fns(egno, ndim, numerical_L_ind):
  """set_fns.py:52-166.  egno 1: L=|alp|^2/2, egno 2: L=indicator{|alp|<=1}, egno 3: Newton (2-D, n_ctrl=1)."""
  print('egno: ', egno, flush=True)
  if numerical_L_ind != 0:
    raise ValueError("ind {} not implemented".format(numerical_L_ind))
  if egno not in (1, 2, 3) or ndim not in (1, 2) or (egno == 3 and ndim != 2):
    raise ValueError("egno {} not implemented".format(egno))
  a = lambda v: (v - 1.0) ** 2 + 0.1
  if egno == 3:
    n_ctrl = 1
    f_fn = lambda alp, x_arr, t_arr: np.concatenate([alp, np.broadcast_to(x_arr[..., 0:1], alp.shape)], axis=-1)
    L_fn = lambda alp, x_arr, t_arr: alp[0][..., 0] ** 2 / 2 + alp[1][..., 0] ** 2 / 2
  elif ndim == 2:
    n_ctrl = 2
    f_fn = lambda alp, x_arr, t_arr: -np.stack([a(x_arr[..., 0]) * alp[..., 0], a(x_arr[..., 1]) * alp[..., 1]], axis=-1)
    L_fn = (lambda alp, x_arr, t_arr: sum(np.sum(al ** 2, axis=-1) / 2 for al in alp)) if egno == 1 else \
           (lambda alp, x_arr, t_arr: 0.0 * alp[0][..., 0])
  else:
    n_ctrl = 1
    f_fn = lambda alp, x_arr, t_arr: -alp * a(x_arr)
    L_fn = (lambda alp, x_arr, t_arr: alp[0][..., 0] ** 2 / 2 + alp[1][..., 0] ** 2 / 2) if egno == 1 else \
           (lambda alp, x_arr, t_arr: 0.0 * alp[0][..., 0])

  def alp_update_fn(*args, **kwargs):
    raise NotImplementedError("the alp proximal step runs inside the CUDA dual-sweep kernel (pdhg_update_dual); "
                              "use update_fns_in_pdhg.update_dual_alternative")
  return Functions(f_fn=f_fn, numerical_L_fn=L_fn, alp_update_fn=alp_update_fn, egno=egno, ndim=ndim, n_ctrl=n_ctrl)

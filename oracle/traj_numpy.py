"""TEST INFRASTRUCTURE ONLY.  NumPy / SciPy restatement of the reference's closed-loop trajectory simulation
(jaxsrc/run_example.py:18-155: compute_traj_1d, extend_bdry_2d, compute_traj_2d), with the Gaussian increments as an explicit
input instead of numpy.random's global stream (`noise[k]` = the draw of step k, shape of x_curr).  Checked against the
reference's own source by oracle/make_golden_traj.py; never imported by the product path.
"""
import numpy as np
from scipy import interpolate

from .pdhg_numpy import get_f_vals


def compute_traj_1d(x_init, alp, f_fn, nt, x_arr, t_arr, x_period, T, epsl=0.0, interp_method='linear', noise=None):
  """run_example.py:18-51.  alp [2, nt-1, nx], x_init [n_sample] -> traj_alp [nt-1, n_sample, 1], traj_x [nt, n_sample]."""
  traj_alp, traj_x = [], [np.asarray(x_init, dtype=float)]
  x_curr = traj_x[0]
  for i in range(nt - 1):
    dt = t_arr[i + 1] - t_arr[i]
    if interp_method == 'linear':
      alp_1 = np.interp(x_curr, x_arr, alp[0, i, :], period=x_period)[:, None]      # :35
      alp_2 = np.interp(x_curr, x_arr, alp[1, i, :], period=x_period)[:, None]      # :36
    elif interp_method == 'nearest':
      idx = np.abs(x_arr - (x_curr % x_period)[..., None]).argmin(axis=-1)            # :38-39
      alp_1, alp_2 = alp[0, i, idx][:, None], alp[1, i, idx][:, None]
    else:
      raise NotImplementedError
    traj_alp.append(alp_1 + alp_2)
    f1, f2 = get_f_vals(f_fn, (alp_1, alp_2), x_curr[:, None] % x_period, T - t_arr[i])   # :46
    vel = f1 + f2
    z = 0.0 if noise is None else noise[i]
    x_curr = x_curr + vel * dt + np.sqrt(2 * epsl * dt) * z                           # :48
    traj_x.append(x_curr)
  return np.stack(traj_alp, axis=0), np.stack(traj_x, axis=0)


def extend_bdry_2d(x_arr, x_min, x_max, val_arr, period, axis, bc, center=False):
  """run_example.py:53-110: periodic copies (bc 0) or repeated edge values (bc 1; zeros for bc 2) so that every sample lies
  inside the interpolation grid; one closing node is appended."""
  if center:
    lb, ub = int(np.floor(x_min / period + 0.5)), int(np.floor(x_max / period + 0.5))
  else:
    lb, ub = int(np.floor(x_min / period)), int(np.floor(x_max / period))
  lb, ub = min(lb, 0), max(ub, 0)
  n_period = ub - lb + 1
  take = lambda a, sl: a[:, sl, :, :] if axis == 1 else a[:, :, sl, :]
  if bc == 0:
    val_arr = np.concatenate([val_arr] * n_period, axis=axis)
    val_arr = np.concatenate([val_arr, take(val_arr, slice(0, 1))], axis=axis)
  else:
    num = val_arr.shape[axis]
    left, right = take(val_arr, slice(0, 1)), take(val_arr, slice(-1, None))
    if bc == 2:
      left, right = np.zeros_like(left), np.zeros_like(right)
    if lb < 0:
      val_arr = np.concatenate([left] * (-lb) * num + [val_arr], axis=axis)
    if ub > 0:
      val_arr = np.concatenate([val_arr] + [right] * (ub * num), axis=axis)
    val_arr = np.concatenate([val_arr, right], axis=axis)
  x_new = np.stack([x_arr] * n_period, axis=0)
  x_new = x_new + np.arange(lb, ub + 1)[:, None] * period
  x_new = np.reshape(x_new, (-1,))
  x_new = np.concatenate([x_new, x_new[0:1] + period * n_period], axis=0)
  return x_new, val_arr


def compute_traj_2d(x_init, alp, f_fn, nt, x1_arr, x2_arr, t_arr, x_period, y_period, T, bc, center, epsl=0.0, interp_method='linear',
                    noise=None):
  """run_example.py:113-155.  alp [4, nt-1, nx, ny, n_ctrl], x_init [n_sample, 2] -> traj_alp [nt-1, n_sample, n_ctrl],
  traj_x [nt, n_sample, 2]."""
  x_init, x1_arr, x2_arr, alp = np.array(x_init, dtype=float), np.array(x1_arr), np.array(x2_arr), np.array(alp)
  traj_alp, traj_x = [], [x_init]
  x_curr = x_init
  (bc_x, bc_y), (cen_x, cen_y) = bc, center
  for i in range(nt - 1):
    dt = t_arr[i + 1] - t_arr[i]
    mn, mx = np.min(x_curr, axis=0), np.max(x_curr, axis=0)
    g1, a_c = extend_bdry_2d(x1_arr, mn[0], mx[0], alp[:, i, :, :, :], x_period, bc=bc_x, axis=1, center=cen_x)
    g2, a_c = extend_bdry_2d(x2_arr, mn[1], mx[1], a_c, y_period, bc=bc_y, axis=2, center=cen_y)
    a = [interpolate.interpn((g1, g2), a_c[j], x_curr, method=interp_method) for j in range(4)]       # :137-140
    traj_alp.append(a[0] + a[1] + a[2] + a[3])
    if bc_x == 0 and bc_y == 0:
      x_in = x_curr % np.array([x_period, y_period])
    elif bc_x == 1 and bc_y == 0:
      x_in = np.array([x_curr[:, 0], x_curr[:, 1] % y_period]).T
    else:
      raise NotImplementedError
    f1_x, f2_x, f1_y, f2_y = get_f_vals(f_fn, tuple(a), x_in, T - t_arr[i])
    vel = np.array([f1_x + f2_x, f1_y + f2_y]).T
    z = 0.0 if noise is None else noise[i]
    x_curr = x_curr + vel * dt + np.sqrt(2 * epsl * dt) * z
    traj_x.append(x_curr)
  return np.stack(traj_alp, axis=0), np.stack(traj_x, axis=0)

"""TEST INFRASTRUCTURE ONLY — the parity oracle.  Never imported by the product path
(`pdhg-optimal-control_b200/`); only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may use it.

NumPy fp64 restatement of the PDHG hot path of TingweiMeng/PDHG-optimal-control.
Citations are relative to `/root/reference/`.

Parity status: the reference ships no tests or golden vectors, and jax/jaxlib
cannot be installed here, so parity is pinned two ways instead:
  * `oracle/make_golden.py` executes the UNMODIFIED reference sources over the NumPy-backed
    API stand-in `oracle/jax_shim.py` and checks this restatement against them operator by
    operator and over whole solves (results committed under `tests/golden/`);
  * the independent known-answer tests T1-T5 of SURVEY.md §4 in `tests/test_oracle_*.py`.
XLA's own rounding (jaxlib FFT, fusion) is not pinned: "parity pinned to the reference source
under NumPy arithmetic, unpinned to jaxlib".

Array conventions are the reference's: phi [K+1, nx(,ny)], rho [K, nx(,ny)], alp = tuple of
2 (1-D) or 4 (2-D) arrays [K, nx(,ny), n_ctrl]; x_arr [1,nx,1] or [1,nx,ny,2]; t slowest.
"""
from collections import namedtuple

import numpy as np
import scipy.fft as sfft

RHO_OFFSET = 1e-4   # update_fns_in_pdhg.py:74,86 ; set_fns.py:104,126,151


# --------------------------------------------------------------------------------------
# Finite-difference stencils  (jaxsrc/utils/utils_diff_op.py, all of it)
# bc: 0 periodic, 1 Neumann, 2 Dirichlet (utils_diff_op.py:5-7).  `axis` 1 = x, 2 = y.
# --------------------------------------------------------------------------------------
def _edge(u, axis, first):
  idx = [slice(None)] * u.ndim
  idx[axis] = slice(0, 1) if first else slice(-1, None)
  return u[tuple(idx)]


def _drop(u, axis, first):
  idx = [slice(None)] * u.ndim
  idx[axis] = slice(1, None) if first else slice(None, -1)
  return u[tuple(idx)]


def diff_right(u, h, axis, bc):
  """(u[i+1]-u[i])/h   utils_diff_op.py:9-23 (x), :93-107 (y)."""
  if bc == 0:
    out = np.roll(u, -1, axis=axis) - u
  elif bc == 1:
    out = np.concatenate([_drop(u, axis, True) - _drop(u, axis, False), np.zeros_like(_edge(u, axis, True))], axis=axis)
  elif bc == 2:
    out = np.concatenate([_drop(u, axis, True), np.zeros_like(_edge(u, axis, True))], axis=axis) - u
  else:
    raise NotImplementedError(bc)
  return out / h


def diff_left(u, h, axis, bc):
  """(u[i]-u[i-1])/h   utils_diff_op.py:51-65 (x), :135-149 (y)."""
  if bc == 0:
    out = u - np.roll(u, 1, axis=axis)
  elif bc == 1:
    out = np.concatenate([np.zeros_like(_edge(u, axis, True)), _drop(u, axis, True) - _drop(u, axis, False)], axis=axis)
  elif bc == 2:
    out = u - np.concatenate([np.zeros_like(_edge(u, axis, True)), _drop(u, axis, False)], axis=axis)
  else:
    raise NotImplementedError(bc)
  return out / h


def diff2(u, h, axis, bc):
  """(u[i+1]+u[i-1]-2u[i])/h**2   utils_diff_op.py:208-226 (x), :255-273 (y)."""
  if bc == 0:
    up = np.roll(u, -1, axis=axis)
    um = np.roll(u, 1, axis=axis)
  elif bc == 1:
    up = np.concatenate([_drop(u, axis, True), _edge(u, axis, False)], axis=axis)
    um = np.concatenate([_edge(u, axis, True), _drop(u, axis, False)], axis=axis)
  elif bc == 2:
    z = np.zeros_like(_edge(u, axis, True))
    up = np.concatenate([_drop(u, axis, True), z], axis=axis)
    um = np.concatenate([z, _drop(u, axis, False)], axis=axis)
  else:
    raise NotImplementedError(bc)
  return (up + um - 2 * u) / h ** 2


def _dec(v):
  """`*_decreasedim`: operator applied to every phi row, rows 1: returned (utils_diff_op.py:34-35 etc.)."""
  return v[1:, ...]


def _inc(v):
  """`*_increasedim`: operator applied to the K dual rows, one zero row prepended (utils_diff_op.py:47-49 etc.)."""
  return np.concatenate([np.zeros_like(v[0:1, ...]), v], axis=0)


def Dx_right_decreasedim(phi, dx, bc): return _dec(diff_right(phi, dx, 1, bc))
def Dx_left_decreasedim(phi, dx, bc): return _dec(diff_left(phi, dx, 1, bc))
def Dy_right_decreasedim(phi, dy, bc): return _dec(diff_right(phi, dy, 2, bc))
def Dy_left_decreasedim(phi, dy, bc): return _dec(diff_left(phi, dy, 2, bc))
def Dxx_decreasedim(phi, dx, bc): return _dec(diff2(phi, dx, 1, bc))
def Dyy_decreasedim(phi, dy, bc): return _dec(diff2(phi, dy, 2, bc))
def Dx_right_increasedim(m, dx, bc): return _inc(diff_right(m, dx, 1, bc))
def Dx_left_increasedim(m, dx, bc): return _inc(diff_left(m, dx, 1, bc))
def Dy_right_increasedim(m, dy, bc): return _inc(diff_right(m, dy, 2, bc))
def Dy_left_increasedim(m, dy, bc): return _inc(diff_left(m, dy, 2, bc))
def Dxx_increasedim(rho, dx, bc): return _inc(diff2(rho, dx, 1, bc))
def Dyy_increasedim(rho, dy, bc): return _inc(diff2(rho, dy, 2, bc))


def Dt_decreasedim(phi, dt):
  """(phi[k+1]-phi[k])/dt   utils_diff_op.py:179-191."""
  return (phi[1:, ...] - phi[:-1, ...]) / dt


def Dt_increasedim(rho, dt):
  """(rho[k]-rho[k-1])/dt with rho[-1]=rho[K]=0   utils_diff_op.py:193-206."""
  z = np.zeros_like(rho[0:1, ...])
  return (np.concatenate([rho, z], axis=0) - np.concatenate([z, rho], axis=0)) / dt


# --------------------------------------------------------------------------------------
# Preconditioner  (jaxsrc/utils/utils_precond.py:10-71,105-178)
# --------------------------------------------------------------------------------------
def tridiagonal_solve(dl, d, du, b):
  """Thomas recurrences of utils_precond.py:10-35 along axis 0; `d`, `b` may carry trailing batch axes."""
  n = d.shape[0]
  tu = np.empty_like(d, dtype=np.result_type(d, du))
  bb = np.empty_like(b, dtype=np.result_type(b, d))
  ex = (slice(None),) + (None,) * (d.ndim - 1)
  dl_ = np.asarray(dl)[ex]
  du_ = np.asarray(du)[ex]
  # forward pass: tu[0] = du[0]/d[0]; tu[i] = du[i]/(d[i]-dl[i]*tu[i-1])  (:13,19-22)
  # (the scan re-applies fwd1 to row 0 with carry du[0]/d[0]; dl[0] = 0 so it is a no-op)
  tu[0] = du_[0] / d[0]
  for i in range(1, n):
    tu[i] = du_[i] / (d[i] - dl_[i] * tu[i - 1])
  # b_[0] = b[0]/d[0]; b_[i] = (b[i]-dl[i]*b_[i-1])/(d[i]-dl[i]*tu[i-1])  (:14,24-27)
  bb[0] = b[0] / d[0]
  for i in range(1, n):
    bb[i] = (b[i] - dl_[i] * bb[i - 1]) / (d[i] - dl_[i] * tu[i - 1])
  # back-substitution x[n-1] = b_[n-1]; x[i] = b_[i]-tu[i]*x[i+1]  (:15,30-33)
  x = np.empty_like(bb)
  x[n - 1] = bb[n - 1]
  for i in range(n - 2, -1, -1):
    x[i] = bb[i] - tu[i] * x[i + 1]
  return x


def compute_Dxx_fft_fv(ndim, nspatial, dspatial, bc):
  """Symbol of the discrete Laplacian, utils_precond.py:42-71."""
  if ndim == 1:
    dx = dspatial[0]
    nx = nspatial[0]
    lap = np.array([-2 / (dx * dx), 1 / (dx * dx)] + [0.0] * (nx - 3) + [1 / (dx * dx)])
    if bc == 0:
      return np.fft.fft(lap)
    if bc == 1:
      return sfft.dct(lap)
    raise NotImplementedError
  if ndim == 2:
    dx, dy = dspatial
    nx, ny = nspatial
    bc_x, bc_y = bc
    lap = np.zeros((nx, ny))
    lap[0, 0] = -2 / (dx * dx) - 2 / (dy * dy)
    lap[0, 1] = 1 / (dy * dy)
    lap[0, -1] = 1 / (dy * dy)
    lap[1, 0] = 1 / (dx * dx)
    lap[-1, 0] = 1 / (dx * dx)
    if bc_x == 0 and bc_y == 0:
      return np.fft.fft2(lap)
    if bc_x == 1 and bc_y == 0:
      return np.fft.fft(sfft.dct(lap, axis=0), axis=-1)
    raise NotImplementedError
  raise NotImplementedError


def _lap_t_diag(K, dt):
  """[2,...,2,1]/dt^2 : Dirichlet-0 before row 0, Neumann after the last row (utils_precond.py:130,166)."""
  return -np.array([-2 / (dt * dt)] * (K - 1) + [-1 / (dt * dt)])


def H1_precond_1d(source_term, fv, dt, bc, C=1.0, pow=1, Ct=1):
  """Solve ((C-Dxx)^pow - Ct*Dtt) u = src[1:], utils_precond.py:105-140."""
  nt, nx = source_term.shape
  K = nt - 1
  if bc != 0:
    raise NotImplementedError
  v = np.fft.fft(source_term[1:, :], axis=1)
  thomas_b = (np.broadcast_to(-fv, (K, nx)) + C) ** pow
  if Ct != 0:
    off = 1 / (dt * dt) * np.ones((K - 1,))
    dl = -np.concatenate([[0.0], off]).astype(np.complex128) * Ct
    du = -np.concatenate([off, [0.0]]).astype(np.complex128) * Ct
    diag = thomas_b + _lap_t_diag(K, dt)[:, None] * Ct
    sol = tridiagonal_solve(dl, diag, du, v)
  else:
    sol = v / thomas_b
  upd = np.fft.ifft(sol, axis=1).real
  return np.concatenate([np.zeros((1, nx)), upd], axis=0)


def H1_precond_2d(source_term, fv, dt, bc, C=1.0):
  """Solve (C - Dtt - Dxx - Dyy) u = src[1:], utils_precond.py:142-178."""
  nt, nx, ny = source_term.shape
  K = nt - 1
  bc_x, bc_y = bc
  if bc_x == 0 and bc_y == 0:
    v = np.fft.fft2(source_term[1:, ...], axes=(1, 2))
  elif bc_x == 1 and bc_y == 0:
    v = np.fft.fft(sfft.dct(source_term[1:, ...], axis=1), axis=2)
  else:
    raise NotImplementedError
  off = 1 / (dt * dt) * np.ones((K - 1,))
  dl = -np.concatenate([[0.0], off]).astype(np.complex128)
  du = -np.concatenate([off, [0.0]]).astype(np.complex128)
  diag = np.broadcast_to(-fv, (K, nx, ny)) + _lap_t_diag(K, dt)[:, None, None] + C
  sol = tridiagonal_solve(dl, diag, du, v)
  if bc_x == 0 and bc_y == 0:
    upd = np.fft.ifft2(sol, axes=(1, 2)).real
  else:
    upd = sfft.idct(np.fft.ifft(sol, axis=2).real, axis=1)
  return np.concatenate([np.zeros((1, nx, ny)), upd], axis=0)


# --------------------------------------------------------------------------------------
# Problem definitions  (jaxsrc/set_fns.py)
# --------------------------------------------------------------------------------------
Functions = namedtuple('Functions', ['f_fn', 'numerical_L_fn', 'alp_update_fn'])   # set_fns.py:164


def set_up_J(egno, ndim, period_spatial):
  """Initial data g, set_fns.py:10-24."""
  if egno != 3:
    if ndim == 1:
      alpha = 2 * np.pi / period_spatial[0]
    elif ndim == 2:
      alpha = np.array([2 * np.pi / period_spatial[0], 2 * np.pi / period_spatial[1]])
    else:
      raise ValueError("ndim {} not implemented".format(ndim))
    return lambda x: np.sum(np.sin(alpha * x), axis=-1)
  x_period, y_period = period_spatial
  return lambda x: np.sin(2 * np.pi / y_period * x[..., 1]) * np.exp(-x[..., 0] ** 2 / 2)


def _coeff_a(x):
  """a(x) = (x-1)^2 + 0.1, set_fns.py:117-118,145."""
  return (x - 1.0) ** 2 + 0.1


def _alp_base(egno, alp_prev, Dphi, param_inv, coeff_f_neg, coeff_H):
  """alp prox without the upwind mask: set_fns.py:63-77 (egno 1), :79-95 (egno 2)."""
  if egno == 1:
    return (Dphi[..., None] * coeff_f_neg + param_inv * alp_prev) / (1 / coeff_H + param_inv)
  if egno == 2:
    nxt = Dphi[..., None] * coeff_f_neg / param_inv + alp_prev
    return np.minimum(coeff_H, np.maximum(-coeff_H, nxt))
  raise ValueError(egno)


def set_up_example_fns(egno, ndim, numerical_L_ind=0):
  """set_fns.py:52-166.  Returns Functions(f_fn, numerical_L_fn, alp_update_fn)."""
  if numerical_L_ind != 0:
    raise ValueError("ind {} not implemented".format(numerical_L_ind))
  if egno == 3:
    # Newton: x=(velocity, position), f=[alp, x_0], L=|alp|^2/2, n_ctrl=1 (set_fns.py:96-111)
    f_fn = lambda alp, x_arr, t_arr: np.concatenate([alp, x_arr[..., 0:1]], axis=-1)
    L1 = lambda alp: alp[..., 0] ** 2 / 1.0 / 2
    numerical_L_fn = lambda alp, x_arr, t_arr: L1(alp[0]) + L1(alp[1])      # n_ctrl == 1 branch, :37-39

    def alp_update_fn(alp_prev, Dphi, rho, sigma, x_arr, t_arr):
      a1x, a2x, a1y, a2y = alp_prev
      Dxr, Dxl, _, _ = Dphi
      p = (rho[..., None] + RHO_OFFSET) / sigma
      coeff_L = 1.0
      n1 = (-Dxr[..., None] + p * a1x) / (coeff_L + p)
      n1 = n1 * (f_fn(n1, x_arr, t_arr)[..., 0:1] >= 0.0)
      n2 = (-Dxl[..., None] + p * a2x) / (coeff_L + p)
      n2 = n2 * (f_fn(n2, x_arr, t_arr)[..., 0:1] < 0.0)
      return (n1, n2, a1y, a2y)
  elif ndim == 2:
    # f_x = -a(x) alp_0, f_y = -a(y) alp_1 (set_fns.py:112-139)
    cf1 = lambda x_arr: np.concatenate([_coeff_a(x_arr[..., 0:1]), np.zeros_like(x_arr[..., 0:1])], axis=-1)
    cf2 = lambda x_arr: np.concatenate([np.zeros_like(x_arr[..., 0:1]), _coeff_a(x_arr[..., 1:2])], axis=-1)
    f_fn = lambda alp, x_arr, t_arr: -np.concatenate(
      [np.sum(cf1(x_arr) * alp, axis=-1, keepdims=True), np.sum(cf2(x_arr) * alp, axis=-1, keepdims=True)], axis=-1)
    if egno != 2:
      L2 = lambda alp, x_arr: np.sum(alp ** 2 / np.ones_like(x_arr), axis=-1) / 2    # set_fns.py:33
    else:
      L2 = lambda alp, x_arr: 0.0 * alp[..., 0]                                     # :36
    numerical_L_fn = lambda alp, x_arr, t_arr: L2(alp[0], x_arr) + L2(alp[1], x_arr) + L2(alp[2], x_arr) + L2(alp[3], x_arr)

    def alp_update_fn(alp_prev, Dphi, rho, sigma, x_arr, t_arr):
      a1x, a2x, a1y, a2y = alp_prev
      Dxr, Dxl, Dyr, Dyl = Dphi
      p = (rho[..., None] + RHO_OFFSET) / sigma
      c1, c2, cH = cf1(x_arr), cf2(x_arr), np.ones_like(x_arr)
      n1x = _alp_base(egno, a1x, Dxr, p, c1, cH)
      n1x = n1x * (f_fn(n1x, x_arr, t_arr)[..., 0:1] >= 0.0)
      n2x = _alp_base(egno, a2x, Dxl, p, c1, cH)
      n2x = n2x * (f_fn(n2x, x_arr, t_arr)[..., 0:1] < 0.0)
      n1y = _alp_base(egno, a1y, Dyr, p, c2, cH)
      n1y = n1y * (f_fn(n1y, x_arr, t_arr)[..., 1:2] >= 0.0)
      n2y = _alp_base(egno, a2y, Dyl, p, c2, cH)
      n2y = n2y * (f_fn(n2y, x_arr, t_arr)[..., 1:2] < 0.0)
      return (n1x, n2x, n1y, n2y)
  elif ndim == 1:
    # f = -a(x) alp (set_fns.py:140-160)
    f_fn = lambda alp, x_arr, t_arr: -alp * _coeff_a(x_arr)
    if egno != 2:
      L1 = lambda alp, x_arr: alp[..., 0] ** 2 / np.ones_like(x_arr)[..., 0] / 2     # set_fns.py:32
    else:
      L1 = lambda alp, x_arr: 0.0 * alp[..., 0]
    numerical_L_fn = lambda alp, x_arr, t_arr: L1(alp[0], x_arr) + L1(alp[1], x_arr)

    def alp_update_fn(alp_prev, Dx_right_phi, Dx_left_phi, rho, sigma, x_arr, t_arr):
      a1, a2 = alp_prev
      p = ((rho + RHO_OFFSET) / sigma)[..., None]
      cf, cH = _coeff_a(x_arr), np.ones_like(x_arr)
      n1 = _alp_base(egno, a1, Dx_right_phi, p, cf, cH)
      n1 = n1 * (f_fn(n1, x_arr, t_arr) >= 0.0)
      n2 = _alp_base(egno, a2, Dx_left_phi, p, cf, cH)
      n2 = n2 * (f_fn(n2, x_arr, t_arr) < 0.0)
      return (n1, n2)
  else:
    raise ValueError("egno {} not implemented".format(egno))
  return Functions(f_fn=f_fn, numerical_L_fn=numerical_L_fn, alp_update_fn=alp_update_fn)


# --------------------------------------------------------------------------------------
# Update operators  (jaxsrc/update_fns_in_pdhg.py)
# --------------------------------------------------------------------------------------
def get_f_vals(f_fn, alp, x_arr, t_arr):
  """Upwind split, update_fns_in_pdhg.py:13-47.  1-D: (f1,f2); 2-D: (f1_x,f2_x,f1_y,f2_y)."""
  out = []
  for j, a in enumerate(alp):
    comp = 0 if j < 2 else 1
    f = f_fn(a, x_arr, t_arr)[..., comp]
    out.append(f * (f >= 0.0) if j % 2 == 0 else f * (f < 0.0))
  return tuple(out)


def _as_bc2(bc, ndim):
  return (bc,) if ndim == 1 else tuple(bc)


def compute_HJ_residual(phi, alp, dt, dspatial, fns_dict, epsl, x_arr, t_arr, bc):
  """update_fns_in_pdhg.py:49-70."""
  ndim = len(dspatial)
  bcs = _as_bc2(bc, ndim)
  L_val = fns_dict.numerical_L_fn(alp, x_arr, t_arr)
  fs = get_f_vals(fns_dict.f_fn, alp, x_arr, t_arr)
  vec = Dt_decreasedim(phi, dt)
  for d in range(ndim):
    vec = vec - epsl * _dec(diff2(phi, dspatial[d], d + 1, bcs[d]))
  adv = 0.0
  for d in range(ndim):
    adv = adv + _dec(diff_right(phi, dspatial[d], d + 1, bcs[d])) * fs[2 * d] \
              + _dec(diff_left(phi, dspatial[d], d + 1, bcs[d])) * fs[2 * d + 1]
  vec = vec - adv
  vec = vec - L_val
  return vec


def compute_cont_residual(rho, alp, dt, dspatial, fns_dict, c_on_rho, epsl, x_arr, t_arr, bc):
  """update_fns_in_pdhg.py:72-96."""
  ndim = len(dspatial)
  bcs = _as_bc2(bc, ndim)
  fs = get_f_vals(fns_dict.f_fn, alp, x_arr, t_arr)
  ms = [(rho + RHO_OFFSET) * f for f in fs]
  res = Dt_increasedim(rho, dt)
  for d in range(ndim):
    res = res + epsl * _inc(diff2(rho, dspatial[d], d + 1, bcs[d]))
  flux = 0.0
  for d in range(ndim):
    flux = flux + _inc(diff_left(ms[2 * d], dspatial[d], d + 1, bcs[d])) \
                + _inc(diff_right(ms[2 * d + 1], dspatial[d], d + 1, bcs[d]))
  res = res - flux
  return np.concatenate([res[:-1, ...], res[-1:, ...] + c_on_rho / dt], axis=0)


def update_rho(rho_prev, phi, alp, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr, bc):
  """update_fns_in_pdhg.py:99-103,115-119."""
  vec = compute_HJ_residual(phi, alp, dt, dspatial, fns_dict, epsl, x_arr, t_arr, bc)
  return np.maximum(rho_prev + sigma * vec, 0.0)


def update_alp(alp_prev, phi, rho, sigma, dspatial, fns_dict, x_arr, t_arr, bc):
  """update_fns_in_pdhg.py:105-113,121-133."""
  ndim = len(dspatial)
  bcs = _as_bc2(bc, ndim)
  if ndim == 1:
    return fns_dict.alp_update_fn(alp_prev, Dx_right_decreasedim(phi, dspatial[0], bcs[0]),
                                  Dx_left_decreasedim(phi, dspatial[0], bcs[0]), rho, sigma, x_arr, t_arr)
  Dphi = (Dx_right_decreasedim(phi, dspatial[0], bcs[0]), Dx_left_decreasedim(phi, dspatial[0], bcs[0]),
          Dy_right_decreasedim(phi, dspatial[1], bcs[1]), Dy_left_decreasedim(phi, dspatial[1], bcs[1]))
  return fns_dict.alp_update_fn(alp_prev, Dphi, rho, sigma, x_arr, t_arr)


def update_primal_1d(phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc,
                     C=1.0, pow=1, Ct=1):
  """update_fns_in_pdhg.py:135-140."""
  delta_phi = compute_cont_residual(rho_prev, alp_prev, dt, dspatial, fns_dict, c_on_rho, epsl, x_arr, t_arr, bc)
  return phi_prev + tau * H1_precond_1d(delta_phi, fv, dt, bc, C=C, pow=pow, Ct=Ct)


def update_primal_2d(phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc,
                     C=1.0, pow=1, Ct=1):
  """update_fns_in_pdhg.py:142-147 (pow, Ct ignored in 2-D)."""
  delta_phi = compute_cont_residual(rho_prev, alp_prev, dt, dspatial, fns_dict, c_on_rho, epsl, x_arr, t_arr, bc)
  return phi_prev + tau * H1_precond_2d(delta_phi, fv, dt, bc, C=C)


def update_dual_oneiter(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, x_arr, t_arr, bc, fns_dict, ndim):
  """update_fns_in_pdhg.py:150-165."""
  alp_next = update_alp(alp_prev, phi_bar, rho_prev, sigma, dspatial, fns_dict, x_arr, t_arr, bc)
  rho_next = update_rho(rho_prev, phi_bar, alp_next, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr, bc)
  with np.errstate(all='ignore'):
    err = np.sum((rho_next - rho_prev) ** 2) / np.sum(rho_next ** 2)
    for alp_p, alp_n in zip(alp_prev, alp_next):
      err = err + np.sum((alp_n - alp_p) ** 2) / np.sum(alp_n ** 2)
  return rho_next, alp_next, err


def update_dual_alternative(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr,
                            ndim, bc, rho_alp_iters=10, eps=1e-7, stats=None):
  """update_fns_in_pdhg.py:167-180.  `stats['n_inner']` (oracle extra) accumulates executed sweeps."""
  n = 0
  for j in range(rho_alp_iters):
    rho_next, alp_next, err = update_dual_oneiter(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl,
                                                  x_arr, t_arr, bc, fns_dict, ndim)
    n += 1
    if err < eps:
      break
    rho_prev = rho_next
    alp_prev = alp_next
  if stats is not None:
    stats['n_inner'] = stats.get('n_inner', 0) + n
    stats.setdefault('inner_hist', {})
    stats['inner_hist'][n] = stats['inner_hist'].get(n, 0) + 1
  return rho_next, alp_next


# --------------------------------------------------------------------------------------
# PDHG loop and time-block marching  (jaxsrc/utils/utils_pdhg_solver.py)
# --------------------------------------------------------------------------------------
def _norm(a):
  return np.sqrt(np.sum(np.asarray(a) ** 2))


def PDHG_solver_oneiter(fn_update_primal, fn_update_dual, fns_dict, phi0, rho0, alp0, x_arr, t_arr,
                        ndim, dt, dspatial, c_on_rho, epsl=0.0, stepsz_param=0.9, fv=None,
                        N_maxiter=1000000, print_freq=1000, eps=1e-6, tfboard=False, tfrecord_ind=0, verbose=False):
  """utils_pdhg_solver.py:9-94."""
  phi_prev, rho_prev, alp_prev = phi0, rho0, alp0
  scale = 1.5
  tau_phi = stepsz_param / scale
  tau_rho = stepsz_param * scale
  error_all, results_all = [], []
  with np.errstate(all='ignore'):
    for i in range(N_maxiter):
      phi_next = fn_update_primal(phi_prev, rho_prev, c_on_rho, alp_prev, tau_phi, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr)
      phi_bar = 2 * phi_next - phi_prev
      rho_next, alp_next = fn_update_dual(phi_bar, rho_prev, c_on_rho, alp_prev, tau_rho, dt, dspatial, epsl,
                                          fns_dict, x_arr, t_arr, ndim, eps=eps)
      err1 = _norm(phi_next - phi_prev) / _norm(phi_prev)
      err2 = _norm(rho_next - rho_prev) / _norm(rho_prev)
      for alp_p, alp_n in zip(alp_prev, alp_next):
        norm_alp = _norm(alp_p)
        norm_err = _norm(alp_p - alp_n)
        if norm_alp < 1e-6 and norm_err > 1e-6:
          err2 += norm_err
        elif norm_alp >= 1e-6:
          err2 += norm_err / norm_alp
      error = np.array([err1, err2])
      if error[0] < eps and error[1] < eps:
        if verbose: print('PDHG converges at iter {}'.format(i), flush=True)
        break
      if np.any(np.isnan(phi_next)) or np.any(np.isnan(rho_next)):
        if verbose: print("Nan error at iter {}".format(i))
        break
      if print_freq > 0 and i % print_freq == 0:
        results_all.append((i, phi_prev, rho_prev, alp_next))
        error_all.append(error)
        if verbose:
          print('iteration {}, primal error {:.2E}, dual error {:.2E}, min rho {:.2f}, max rho {:.2f}'.format(
            i, error[0], error[1], np.min(rho_next), np.max(rho_next)), flush=True)
      phi_prev, rho_prev, alp_prev = phi_next, rho_next, alp_next
  if verbose:
    # the reference prints error[2], which JAX clamps to error[1] (:90)
    print('iteration {}, primal error with prev step {:.2E}, dual error with prev step {:.2E}, eqt error {:.2E}'.format(
      i, error[0], error[1], error[1]), flush=True)
  results_all.append((i + 1, phi_next, rho_next, alp_next))
  error_all.append(error)
  return results_all, np.array(error_all)


def PDHG_multi_step(fn_update_primal, fn_update_dual, fns_dict, g, x_arr,
                    ndim, nt, nspatial, dt, dspatial, c_on_rho, time_step_per_PDHG=2,
                    epsl=0.0, stepsz_param=0.9, n_ctrl=None, fv=None,
                    N_maxiter=1000000, print_freq=1000, eps=1e-6, tfboard=False, verbose=False, info=None):
  """utils_pdhg_solver.py:97-225 (save/load of middle results omitted: out of the hot path).
  `info` (oracle extra): dict that receives per-block iteration counts and the step sizes tried."""
  if n_ctrl is None:
    n_ctrl = ndim
  tsp = time_step_per_PDHG
  assert (nt - 1) % (tsp - 1) == 0
  nt_PDHG = (nt - 1) // (tsp - 1)
  phi0 = np.concatenate([g] * tsp, axis=0)        # einshape "i...->(ki)..." with i = 1 (:123)
  shape_d = (tsp - 1,) + tuple(nspatial)
  rho0 = np.zeros(shape_d) + c_on_rho
  alp0 = tuple(np.zeros(shape_d + (n_ctrl,)) for _ in range(2 * ndim))
  max_iters = 0
  phi_all, rho_all, alp_all, errs_all = [], [], [], []
  stepsz_param_min = stepsz_param / 10
  stepsz_param_delta = stepsz_param / 10
  sol_nan = False
  if info is not None:
    info.update(block_iters=[], stepsz_tried=[], stepsz_used=[])
  for i in range(nt_PDHG):
    t_arr = np.linspace(i * dt * (tsp - 1), (i + 1) * dt * (tsp - 1), num=tsp)[1:]
    t_arr = t_arr[:, None] if ndim == 1 else t_arr[:, None, None]
    while True:
      if info is not None:
        info['stepsz_tried'].append((i, stepsz_param))
      results_all, errs = PDHG_solver_oneiter(fn_update_primal, fn_update_dual, fns_dict, phi0, rho0, alp0, x_arr, t_arr,
                                              ndim, dt, dspatial, c_on_rho, epsl=epsl, stepsz_param=stepsz_param, fv=fv,
                                              N_maxiter=N_maxiter, print_freq=print_freq, eps=eps, verbose=verbose)
      if np.any(np.isnan(errs)):
        if stepsz_param > stepsz_param_min + stepsz_param_delta:
          stepsz_param -= stepsz_param_delta
          if verbose: print('pdhg does not conv at t_ind = {}, decrease step size to {}'.format(i, stepsz_param), flush=True)
        else:
          if verbose: print('pdhg does not conv at t_ind = {}, algorithm failed'.format(i), flush=True)
          sol_nan = True
          break
      else:
        pdhg_iters, phi_curr, rho_curr, alp_curr = results_all[-1]
        max_iters = max(max_iters, pdhg_iters)
        phi_all.append(phi_curr[:-1, :] if i < nt_PDHG - 1 else phi_curr)
        rho_all.append(rho_curr)
        alp_all.append(np.stack(alp_curr, axis=0))
        errs_all.append(errs)
        if info is not None:
          info['block_iters'].append(pdhg_iters)
          info['stepsz_used'].append(stepsz_param)
        g_diff = phi_curr[-1:, ...] - phi0[0:1, ...]
        phi0 = phi0 + g_diff
        rho0 = rho_curr
        alp0 = alp_curr
        break
    if sol_nan:
      break
  if info is not None:
    info['sol_nan'] = sol_nan
    info['stepsz_final'] = stepsz_param
  if len(phi_all) == 0:
    # the reference raises on concatenate([]) here (:215); the oracle reports the failure instead
    return [(max_iters, None, None, None)], errs_all
  phi_out = np.concatenate(phi_all, axis=0)
  rho_out = np.concatenate(rho_all, axis=0)
  alp_out = np.concatenate(alp_all, axis=1)
  return [(max_iters, phi_out, rho_out, alp_out)], errs_all


# --------------------------------------------------------------------------------------
# Driver slice  (jaxsrc/run_example.py:157-210,228-240,273-287)
# --------------------------------------------------------------------------------------
def make_grid(egno, ndim, nx, ny, x_period, y_period):
  """x_arr, bc, n_ctrl as run_example.py:228-240,273-287 builds them."""
  centered = (egno == 3)
  if egno == 3:
    assert ndim == 2
    n_ctrl, bc = 1, (1, 0)
  else:
    n_ctrl, bc = ndim, (0 if ndim == 1 else (0, 0))
  if ndim == 1:
    x_arr = np.linspace(0.0, x_period, num=nx, endpoint=False)[None, :, None]
    if centered:
      x_arr = x_arr - x_period / 2
  else:
    x1 = np.linspace(0.0, x_period, num=nx, endpoint=False)
    x2 = np.linspace(0.0, y_period, num=ny, endpoint=False)
    if centered:
      x1 = x1 - x_period / 2
      x2 = x2 - y_period / 2
    xm, ym = np.meshgrid(x1, x2, indexing='ij')
    x_arr = np.stack([xm, ym], axis=-1)[None, ...]
  return x_arr, bc, n_ctrl


def solve_HJ(ndim, n_ctrl, egno, epsl, fns_dict, nx, ny, nt, x_period, y_period, T, x_arr,
             c_on_rho, time_step_per_PDHG, stepsz_param, N_maxiter, print_freq, eps, bc,
             C=1.0, pow=1.0, Ct=1.0, verbose=False, info=None, g=None, stats=None):
  """run_example.py:157-210 without plotting.  `g` (oracle extra) overrides the initial data."""
  dt = T / (nt - 1)
  dx = x_period / nx
  dy = y_period / ny
  if ndim == 1:
    period_spatial, dspatial, nspatial = (x_period,), (dx,), (nx,)
  else:
    period_spatial, dspatial, nspatial = (x_period, y_period), (dx, dy), (nx, ny)
  if g is None:
    g = set_up_J(egno, ndim, period_spatial)(x_arr)
  fv = compute_Dxx_fft_fv(ndim, nspatial, dspatial, bc)
  up = update_primal_1d if ndim == 1 else update_primal_2d
  fn_update_primal = lambda phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr: \
    up(phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc, C=C, pow=pow, Ct=Ct)
  fn_update_dual = lambda phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr, ndim, eps: \
    update_dual_alternative(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr, ndim, bc,
                            eps=eps, stats=stats)
  return PDHG_multi_step(fn_update_primal, fn_update_dual, fns_dict, g, x_arr, ndim, nt, nspatial, dt, dspatial, c_on_rho,
                         time_step_per_PDHG=time_step_per_PDHG, epsl=epsl, stepsz_param=stepsz_param, fv=fv, n_ctrl=n_ctrl,
                         N_maxiter=N_maxiter, print_freq=print_freq, eps=eps, verbose=verbose, info=info)

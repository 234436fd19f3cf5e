"""CPU: independent known-answer tests that pin the oracle."""
import numpy as np
import pytest

from oracle import pdhg_numpy as orc


def _setup(ndim, nx, ny, K, egno=1, seed=0):
  rng = np.random.default_rng(seed)
  x_arr, bc, n_ctrl = orc.make_grid(egno, ndim, nx, ny, 2.0, 2.0)
  nsp = (nx,) if ndim == 1 else (nx, ny)
  dsp = (2.0 / nx,) if ndim == 1 else (2.0 / nx, 2.0 / ny)
  phi = rng.standard_normal((K + 1,) + nsp)
  rho = np.abs(rng.standard_normal((K,) + nsp)) * 10
  alp = []
  for j in range(2 * ndim):
    a = rng.standard_normal((K,) + nsp + (n_ctrl,))
    if ndim == 2:
      a[..., 1 if j < 2 else 0] = 0.0
    alp.append(a)
  return x_arr, bc, nsp, dsp, phi, rho, tuple(alp)


@pytest.mark.parametrize("ndim,nx,ny,K", [(1, 16, 1, 3), (2, 8, 6, 2)])
def test_T1_adjointness(ndim, nx, ny, K):
  """sum (HJ_residual(phi; alp) + L(alp)) * rho == - sum phi * cont_residual(rho; alp) with the 1e-4 offset and c_on_rho
  removed: the continuity residual is minus the adjoint of the HJ operator (what makes the iteration a PDHG step)."""
  x_arr, bc, nsp, dsp, phi, rho, alp = _setup(ndim, nx, ny, K)
  fns = orc.set_up_example_fns(1, ndim, 0)
  dt, epsl = 0.05, 0.3
  saved = orc.RHO_OFFSET
  orc.RHO_OFFSET = 0.0
  try:
    hj = orc.compute_HJ_residual(phi, alp, dt, dsp, fns, epsl, x_arr, None, bc) + fns.numerical_L_fn(alp, x_arr, None)
    cont = orc.compute_cont_residual(rho, alp, dt, dsp, fns, 0.0, epsl, x_arr, None, bc)
  finally:
    orc.RHO_OFFSET = saved
  lhs, rhs = np.sum(hj * rho), -np.sum(phi * cont)
  assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), 1.0)


@pytest.mark.parametrize("K", [1, 2, 5])
def test_T2_preconditioner_inverse_1d(K):
  nx, dt, dx, C, pw, Ct = 20, 0.1, 0.1, 0.7, 1.0, 1.3
  rng = np.random.default_rng(1)
  src = rng.standard_normal((K + 1, nx))
  fv = orc.compute_Dxx_fft_fv(1, (nx,), (dx,), 0)
  u = orc.H1_precond_1d(src, fv, dt, 0, C=C, pow=pw, Ct=Ct)
  assert np.all(u[0] == 0.0)
  v = u[1:]
  lap_x = orc.diff2(v, dx, 1, 0)
  up = np.concatenate([v[1:], v[-1:]], axis=0)           # Neumann after the last row
  um = np.concatenate([np.zeros((1, nx)), v[:-1]], axis=0)  # Dirichlet-0 before row 0
  dtt = (up + um - 2 * v) / dt ** 2
  back = C * v - lap_x - Ct * dtt
  assert np.max(np.abs(back - src[1:])) < 1e-10 * np.max(np.abs(src))


def test_T2_preconditioner_inverse_2d():
  K, nx, ny, dt = 3, 8, 6, 0.2
  dsp = (2.0 / nx, 2.0 / ny)
  rng = np.random.default_rng(2)
  src = rng.standard_normal((K + 1, nx, ny))
  fv = orc.compute_Dxx_fft_fv(2, (nx, ny), dsp, (0, 0))
  v = orc.H1_precond_2d(src, fv, dt, (0, 0), C=1.0)[1:]
  lap = orc.diff2(v, dsp[0], 1, 0) + orc.diff2(v, dsp[1], 2, 0)
  up = np.concatenate([v[1:], v[-1:]], axis=0)
  um = np.concatenate([np.zeros((1, nx, ny)), v[:-1]], axis=0)
  back = v - lap - (up + um - 2 * v) / dt ** 2
  assert np.max(np.abs(back - src[1:])) < 1e-10 * np.max(np.abs(src))


@pytest.mark.parametrize("n", [1, 2, 5, 64])
def test_T3_thomas_vs_dense(n):
  rng = np.random.default_rng(n)
  d = 4.0 + rng.random(n)
  dl = np.concatenate([[0.0], -rng.random(n - 1)])
  du = np.concatenate([-rng.random(n - 1), [0.0]])
  b = rng.standard_normal(n)
  M = np.diag(d) + np.diag(dl[1:], -1) + np.diag(du[:-1], 1)
  assert np.max(np.abs(orc.tridiagonal_solve(dl, d, du, b) - np.linalg.solve(M, b))) < 1e-13


def test_T4_symbol_closed_form():
  nx, ny, dx, dy = 160, 24, 2.0 / 160, 2.0 / 24
  fv = orc.compute_Dxx_fft_fv(1, (nx,), (dx,), 0)
  k = np.arange(nx)
  assert np.max(np.abs(fv.real - (2 * np.cos(2 * np.pi * k / nx) - 2) / dx ** 2)) < 1e-9 * 4 / dx ** 2
  assert np.max(np.abs(fv.imag)) < 1e-9 * 4 / dx ** 2
  fv2 = orc.compute_Dxx_fft_fv(2, (nx, ny), (dx, dy), (0, 0))
  ref = (2 * np.cos(2 * np.pi * k[:, None] / nx) - 2) / dx ** 2 + (2 * np.cos(2 * np.pi * np.arange(ny)[None] / ny) - 2) / dy ** 2
  assert np.max(np.abs(fv2.real - ref)) < 1e-9 * 4 / dx ** 2


def test_T5_converged_block_solves_implicit_engquist_osher():
  """For egno 1 a converged block satisfies (phi1-phi0)/dt + a^2/2 (min(D+ phi1,0)^2 + max(D- phi1,0)^2) = 0."""
  nx, nt = 40, 11
  x_arr, bc, n_ctrl = orc.make_grid(1, 1, nx, 1, 2.0, 2.0)
  fns = orc.set_up_example_fns(1, 1, 0)
  res, _ = orc.solve_HJ(1, n_ctrl, 1, 0.0, fns, nx, 1, 3, 2.0, 2.0, 2.0 / (nt - 1), x_arr, 70.0, 2, 0.1, 100000, 100000, 1e-6, bc)
  phi = res[0][1]
  dt, dx = 1.0 / (nt - 1), 2.0 / nx
  a = (x_arr[0, :, 0] - 1.0) ** 2 + 0.1
  for k in range(2):
    p1 = phi[k + 1]
    dp = (np.roll(p1, -1) - p1) / dx
    dm = (p1 - np.roll(p1, 1)) / dx
    r = (p1 - phi[k]) / dt + a ** 2 / 2 * (np.minimum(dp, 0) ** 2 + np.maximum(dm, 0) ** 2)
    assert np.max(np.abs(r)) < 5e-3 * max(np.max(np.abs((p1 - phi[k]) / dt)), 1.0)

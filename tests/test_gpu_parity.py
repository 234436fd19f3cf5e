"""GPU: the CUDA path (through the C ABI, via the Python mirror of the reference API) against the golden vectors
generated from the reference sources and against the NumPy oracle on the same inputs.
Parity bar (BASELINE.json north_star): rel-Linf <= 1e-10 in fp64 on phi / rho / alp, identical stopping iteration
per time block, identical step-size fallback sequence."""
import os

import numpy as np
import pytest

from helpers import TOL, dspatial_of, golden, golden_names, quiet, relmax

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pk(built_lib):
  import pdhg_b200
  from pdhg_b200 import run_example, set_fns, update_fns_in_pdhg
  from pdhg_b200.utils import utils_pdhg_solver
  os.environ.pop("PDHG_FORCE_PATH", None)
  return dict(rx=run_example, sf=set_fns, upd=update_fns_in_pdhg, sol=utils_pdhg_solver)


def _problem(pk, d):
  egno, ndim, nx, ny = [int(d[k]) for k in ("egno", "ndim", "nx", "ny")]
  n_ctrl, bc, cen = pk["rx"].problem_setup(egno, ndim)
  x_arr = pk["rx"].make_x_arr(ndim, nx, ny, 2.0, 2.0, cen)
  fns, _ = quiet(pk["sf"].set_up_example_fns, egno, ndim, 0)
  return egno, ndim, nx, ny, n_ctrl, bc, x_arr, fns


OPS = golden_names("op_")


@pytest.mark.parametrize("name", OPS)
def test_update_operators_vs_reference_golden(pk, name):
  d = golden(name)
  egno, ndim, nx, ny, n_ctrl, bc, x_arr, fns = _problem(pk, d)
  upd = pk["upd"]
  dsp, dt, epsl = dspatial_of(d), float(d["dt"]), float(d["epsl"])
  alp = tuple(d["alp"][j] for j in range(2 * ndim))
  fn = upd.update_primal_1d if ndim == 1 else upd.update_primal_2d
  pn = fn(d["phi"], d["rho"], 70.0, alp, float(d["tau"]), dt, dsp, fns, None, epsl, x_arr, None, bc,
          C=float(d["C"]), pow=float(d["pow"]), Ct=float(d["Ct"]))
  assert relmax(pn, d["phi_next"]) < TOL
  r1, a1, e1 = upd.update_dual_oneiter(d["phi_bar"], d["rho"], 70.0, alp, float(d["sigma"]), dt, dsp, epsl, x_arr, None, bc, fns, ndim)
  assert relmax(r1, d["rho_sweep1"]) < TOL and relmax(np.stack(a1), d["alp_sweep1"]) < TOL
  eref = float(d["err_sweep1"])       # NaN for egno 3 (0/0 on the untouched y pair, update_fns_in_pdhg.py:164)
  assert (np.isnan(e1) and np.isnan(eref)) or abs(e1 - eref) <= 1e-9 * abs(eref)
  rN, aN, _, n_inner = upd._update_dual(d["phi_bar"], d["rho"], 70.0, alp, float(d["sigma"]), dt, dsp, epsl, fns, x_arr, None, ndim, bc,
                                        10, float(d["eps"]))
  assert relmax(rN, d["rho_dual"]) < TOL and relmax(np.stack(aN), d["alp_dual"]) < TOL
  assert n_inner == int(d["n_inner"])


SOLVES = [n for n in golden_names("solve_") if "failed" not in n]


@pytest.mark.parametrize("path", [1, 2])
@pytest.mark.parametrize("name", SOLVES)
def test_solve_HJ_vs_reference_golden(pk, name, path):
  d = golden(name)
  egno, ndim, nx, ny, n_ctrl, bc, x_arr, fns = _problem(pk, d)
  if path == 1 and ndim == 2:
    pytest.skip("the single-CTA kernel is 1-D only")
  nt, tsp = int(d["nt"]), int(d["tsp"])
  os.environ["PDHG_FORCE_PATH"] = str(path)
  try:
    info = {}
    (res, errs), log = quiet(pk["rx"].solve_HJ, ndim, n_ctrl, egno, float(d["epsl"]), fns, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, tsp,
                             float(d["stepsz"]), int(d["N_maxiter"]), int(d["print_freq"]), 1e-6, bc,
                             C=float(d["C"]), pow=float(d["pow"]), Ct=float(d["Ct"]), info=info)
  finally:
    os.environ.pop("PDHG_FORCE_PATH", None)
  assert info["path"] == path
  mi, phi, rho, alp = res[0]
  assert info["block_iters"] == d["block_iters"].tolist()            # same stopping iteration in every block
  assert info["stepsz_used"] == d["stepsz_used"].tolist()            # bit-identical fallback arithmetic
  assert int(mi) == int(d["max_iters"]) and info["sol_nan"] == bool(d["sol_nan"])
  assert relmax(phi, d["phi"]) < TOL and relmax(rho, d["rho"]) < TOL and relmax(alp, d["alp"]) < TOL
  assert [len(e) for e in errs] == d["errs_nrec"].tolist()
  assert relmax(np.concatenate([e.reshape(-1, 2) for e in errs]), d["errs_flat"]) < 1e-7
  for s in d["stepsz_decrements"].tolist():                           # the reference's fallback announcements
    assert "decrease step size to {}".format(s) in log


@pytest.mark.parametrize("path", [1, 2])
def test_algorithm_failed_path_vs_reference_golden(pk, path):
  """Every step size 0.1 -> 0.01 NaNs in block 0: the reference announces nine decrements, then 'algorithm failed'
  (utils_pdhg_solver.py:180-187) and crashes; here the same sequence is reported and the failure is a status."""
  d = golden("solve_1d_eg1_nx32_nt5_failed")
  egno, ndim, nx, ny, n_ctrl, bc, x_arr, fns = _problem(pk, d)
  os.environ["PDHG_FORCE_PATH"] = str(path)
  try:
    info = {}
    (res, errs), log = quiet(pk["rx"].solve_HJ, ndim, n_ctrl, egno, float(d["epsl"]), fns, nx, ny, int(d["nt"]), 2.0, 2.0, 1.0, x_arr, 70.0,
                             int(d["tsp"]), float(d["stepsz"]), int(d["N_maxiter"]), int(d["print_freq"]), 1e-6, bc, info=info)
  finally:
    os.environ.pop("PDHG_FORCE_PATH", None)
  assert info["sol_nan"] and info["blocks_done"] == 0 and res[0][1] is None and errs == []
  assert info["stepsz_final"] == d["stepsz_decrements"].tolist()[-1]
  for s in d["stepsz_decrements"].tolist():
    assert "decrease step size to {}".format(s) in log
  assert "algorithm failed" in log


def test_cfg1_readme_example_full_solve(pk):
  """BASELINE configs[0]: egno=1 ndim=1 epsl=0 nx=160 nt=41 stepsz 0.1 to convergence over all 40 blocks
  (131k iterations) against the oracle fixture (scripts/make_oracle_golden.py cfg1)."""
  d = golden("oracle_cfg1")
  egno, ndim, nx, ny, n_ctrl, bc, x_arr, fns = _problem(pk, d)
  info = {}
  (res, errs), _ = quiet(pk["rx"].solve_HJ, 1, n_ctrl, 1, 0.0, fns, 160, 1, 41, 2.0, 2.0, 1.0, x_arr, 70.0, 2, 0.1, 1000000, 10000, 1e-6, bc,
                         info=info)
  mi, phi, rho, alp = res[0]
  assert info["block_iters"] == d["block_iters"].tolist() and sum(info["block_iters"]) == 130929
  assert info["n_inner"] == int(d["n_inner"])
  assert relmax(phi, d["phi"]) < TOL and relmax(rho, d["rho"]) < TOL and relmax(alp, d["alp"]) < TOL
  assert relmax(np.concatenate([e.reshape(-1, 2) for e in errs]), d["errs_flat"]) < 1e-7


@pytest.mark.parametrize("ndim,nx,ny,K,epsl,nmax,pf", [(1, 48, 1, 1, 0.0, 700, 250), (1, 20, 1, 3, 0.02, 300, 100), (2, 12, 10, 2, 0.05, 130, 50)])
def test_PDHG_solver_oneiter_snapshots_vs_oracle(pk, ndim, nx, ny, K, epsl, nmax, pf):
  """results_all snapshots (i, phi_prev, rho_prev, alp_next) at every i % print_freq == 0 and error_all rows."""
  from oracle import pdhg_numpy as orc
  rx, upd, sol = pk["rx"], pk["upd"], pk["sol"]
  n_ctrl, bc, cen = rx.problem_setup(1, ndim)
  x_arr = rx.make_x_arr(ndim, nx, ny, 2.0, 2.0, cen)
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, ndim, 0)
  fns_o = orc.set_up_example_fns(1, ndim, 0)
  nsp = (nx,) if ndim == 1 else (nx, ny)
  dsp = (2.0 / nx,) if ndim == 1 else (2.0 / nx, 2.0 / ny)
  dt = 0.05
  g = pk["sf"].set_up_J(1, ndim, (2.0,) * ndim)(x_arr)
  phi0 = np.concatenate([g] * (K + 1), axis=0)
  rho0 = np.zeros((K,) + nsp) + 70.0
  alp0 = tuple(np.zeros((K,) + nsp + (n_ctrl,)) for _ in range(2 * ndim))
  (res, errs), _ = quiet(sol.PDHG_solver_oneiter, upd.NativeUpdatePrimal(ndim, bc), upd.NativeUpdateDual(bc), fns, phi0, rho0, alp0, x_arr, None,
                         ndim, dt, dsp, 70.0, epsl=epsl, stepsz_param=0.1, fv=None, N_maxiter=nmax, print_freq=pf, eps=1e-6)
  fv = orc.compute_Dxx_fft_fv(ndim, nsp, dsp, bc)
  up_o = orc.update_primal_1d if ndim == 1 else orc.update_primal_2d
  prim = lambda *a: up_o(*a, bc)
  dual = lambda *a, eps: orc.update_dual_alternative(*a, bc, eps=eps)
  res_o, errs_o = orc.PDHG_solver_oneiter(prim, dual, fns_o, phi0, rho0, alp0, x_arr, None, ndim, dt, dsp, 70.0, epsl=epsl, stepsz_param=0.1,
                                          fv=fv, N_maxiter=nmax, print_freq=pf, eps=1e-6)
  assert len(res) == len(res_o) and errs.shape == errs_o.shape
  assert relmax(errs, errs_o) < 1e-7
  for (i, p, r, a), (io, po, ro, ao) in zip(res, res_o):
    assert i == io
    assert relmax(p, po) < TOL and relmax(r, ro) < TOL and relmax(np.stack(a), np.stack(ao)) < TOL


def test_injected_callables_host_loop_equals_fused_loop(pk):
  """The operator-injection boundary: arbitrary callables with the reference signatures drive the host loop;
  with our own operators wrapped in lambdas the result must equal the fused on-device loop bit for bit."""
  rx, upd, sol = pk["rx"], pk["upd"], pk["sol"]
  nx, K = 32, 1
  n_ctrl, bc, _ = rx.problem_setup(1, 1)
  x_arr = rx.make_x_arr(1, nx, 1, 2.0, 2.0)
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, 1, 0)
  g = pk["sf"].set_up_J(1, 1, (2.0,))(x_arr)
  phi0, rho0 = np.concatenate([g] * 2, axis=0), np.zeros((1, nx)) + 70.0
  alp0 = (np.zeros((1, nx, 1)), np.zeros((1, nx, 1)))
  P, D = upd.NativeUpdatePrimal(1, bc), upd.NativeUpdateDual(bc)
  args = (fns, phi0, rho0, alp0, x_arr, None, 1, 0.1, (2.0 / nx,), 70.0)
  kw = dict(epsl=0.01, stepsz_param=0.1, fv=None, N_maxiter=40, print_freq=10, eps=1e-6)
  (res_f, err_f), _ = quiet(sol.PDHG_solver_oneiter, P, D, *args, **kw)
  (res_h, err_h), _ = quiet(sol.PDHG_solver_oneiter, lambda *a: P(*a), lambda *a, eps: D(*a, eps=eps), *args, **kw)
  assert len(res_f) == len(res_h)
  assert relmax(err_h, err_f) < 1e-9
  for (i, p, r, a), (ih, ph, rh, ah) in zip(res_f, res_h):
    assert i == ih and relmax(ph, p) < 1e-12 and relmax(rh, r) < 1e-12 and relmax(np.stack(ah), np.stack(a)) < 1e-12


def test_cfg3_grid_first_iterations_vs_oracle(pk):
  """BASELINE configs[2] grid (2-D 256x256, nt=65, tsp=2): 25 outer iterations of block 0 against the oracle."""
  from oracle import pdhg_numpy as orc
  rx = pk["rx"]
  nx = ny = 256
  n_ctrl, bc, _ = rx.problem_setup(1, 2)
  x_arr = rx.make_x_arr(2, nx, ny, 2.0, 2.0)
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, 2, 0)
  info = {}
  (res, errs), _ = quiet(rx.solve_HJ, 2, n_ctrl, 1, 0.0, fns, nx, ny, 2, 2.0, 2.0, 1.0 / 64, x_arr, 70.0, 2, 0.1, 25, 10, 1e-6, bc, info=info)
  res_o, errs_o = orc.solve_HJ(2, n_ctrl, 1, 0.0, orc.set_up_example_fns(1, 2, 0), nx, ny, 2, 2.0, 2.0, 1.0 / 64, x_arr, 70.0, 2, 0.1, 25, 10,
                               1e-6, bc)
  for a, b in zip(res[0][1:], res_o[0][1:]):
    assert relmax(a, b) < TOL
  assert relmax(errs[0], errs_o[0]) < 1e-7 and info["block_iters"] == [25]


@pytest.mark.parametrize("ndim,nx,ny,nt,epsl,nmax", [(2, 256, 256, 4, 0.002, 6), (2, 64, 256, 3, 0.0, 8), (2, 256, 48, 3, 0.005, 8), (2, 32, 64, 5, 0.01, 12), (1, 256, 1, 4, 0.003, 40),
                                                       (1, 256, 1, 6, 0.0, 30)])
def test_warp_private_256_point_transforms_vs_oracle(pk, ndim, nx, ny, nt, epsl, nmax):
  """The barrier-free 256-point fast path of the cooperative kernel (phases A/C when ny == 256, phase B when nx == 256; 1-D
  grids run with their x axis as the contiguous axis) on coupled space-time blocks (K = nt - 1 > 1, Thomas in t), including
  row counts that leave the last 4-row unit partly empty, against the oracle; and == the generic tiled path (PDHG_NO_W256)."""
  from oracle import pdhg_numpy as orc
  rx = pk["rx"]
  n_ctrl, bc, _ = rx.problem_setup(1, ndim)
  x_arr = rx.make_x_arr(ndim, nx, ny, 2.0, 2.0)
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, ndim, 0)
  out = []
  T = (nt - 1) / 64.0
  os.environ["PDHG_FORCE_PATH"] = "2"
  try:
    # (by default the warp-private transforms are only taken on grids with enough rows per SM)
    for envs in (("PDHG_FORCE_W256",), ("PDHG_NO_W256",)):
      for env in envs:
        os.environ[env] = "1"
      (res, errs), _ = quiet(rx.solve_HJ, ndim, n_ctrl, 1, epsl, fns, nx, ny, nt, 2.0, 2.0, T, x_arr, 70.0, nt, 0.1, nmax, 10, 1e-6, bc)
      for env in envs:
        os.environ.pop(env)
      out.append((res, errs))
  finally:
    for env in ("PDHG_FORCE_PATH", "PDHG_FORCE_W256", "PDHG_NO_W256"):
      os.environ.pop(env, None)
  res_o, errs_o = orc.solve_HJ(ndim, n_ctrl, 1, epsl, orc.set_up_example_fns(1, ndim, 0), nx, ny, nt, 2.0, 2.0, T, x_arr, 70.0, nt, 0.1, nmax, 10,
                               1e-6, bc)
  for res, errs in out:
    for a, b in zip(res[0][1:], res_o[0][1:]):
      assert relmax(a, b) < TOL
    assert relmax(errs[0], errs_o[0]) < 1e-7
  for a, b in zip(out[0][0][0][1:], out[1][0][0][1:]):
    assert relmax(a, b) < 1e-12


@pytest.mark.parametrize("nx,ny", [(2048, 8), (8, 2048), (512, 24)])
def test_long_transform_axes_vs_oracle(pk, nx, ny):
  """BASELINE configs[4] axis length (2048 = 16 * 16 * 8 points, three Stockham stages) along x and along y, viscous, tiny step
  size as SURVEY.md section 8(d) prescribes for that grid: first iterations against the oracle."""
  from oracle import pdhg_numpy as orc
  rx = pk["rx"]
  n_ctrl, bc, _ = rx.problem_setup(1, 2)
  x_arr = rx.make_x_arr(2, nx, ny, 2.0, 2.0)
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, 2, 0)
  (res, errs), _ = quiet(rx.solve_HJ, 2, n_ctrl, 1, 0.1, fns, nx, ny, 2, 2.0, 2.0, 1.0 / 256, x_arr, 70.0, 2, 5e-4, 12, 10, 1e-6, bc)
  res_o, errs_o = orc.solve_HJ(2, n_ctrl, 1, 0.1, orc.set_up_example_fns(1, 2, 0), nx, ny, 2, 2.0, 2.0, 1.0 / 256, x_arr, 70.0, 2, 5e-4, 12, 10,
                               1e-6, bc)
  for a, b in zip(res[0][1:], res_o[0][1:]):
    assert relmax(a, b) < TOL
  assert relmax(errs[0], errs_o[0]) < 1e-7


def test_spacetime_block_2d_vs_oracle(pk):
  """time_step_per_PDHG = nt (one space-time block, the HBM-bound regime): FFT-xy + Thomas-t over K = 8 rows."""
  from oracle import pdhg_numpy as orc
  rx = pk["rx"]
  nx, ny, nt = 64, 48, 9
  n_ctrl, bc, _ = rx.problem_setup(1, 2)
  x_arr = rx.make_x_arr(2, nx, ny, 2.0, 2.0)
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, 2, 0)
  (res, errs), _ = quiet(rx.solve_HJ, 2, n_ctrl, 1, 0.02, fns, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, nt, 0.1, 40, 10, 1e-6, bc)
  res_o, errs_o = orc.solve_HJ(2, n_ctrl, 1, 0.02, orc.set_up_example_fns(1, 2, 0), nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, nt, 0.1, 40, 10, 1e-6, bc)
  for a, b in zip(res[0][1:], res_o[0][1:]):
    assert relmax(a, b) < TOL
  assert relmax(errs[0], errs_o[0]) < 1e-7


def test_batched_instances_equal_single_instance_runs(pk):
  """cfg4-style sweep: B instances with different initial data / epsl in one launch == B separate solves (bitwise),
  and each equals the oracle."""
  from oracle import pdhg_numpy as orc
  rx = pk["rx"]
  nx, nt, B = 64, 5, 6
  rng = np.random.default_rng(0)
  A_, th, u = rng.uniform(0.5, 1.5, B), rng.uniform(0, 2 * np.pi, B), rng.uniform(0, 1, B)
  x_arr = rx.make_x_arr(1, nx, 1, 2.0, 2.0)
  g = A_[:, None] * np.sin(np.pi * x_arr[0, :, 0][None, :] + th[:, None])
  epsl = 0.002 * u
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, 1, 0)
  phi, rho, alp, logs = rx.solve_HJ_batch(1, 1, 1, epsl, fns, nx, 1, nt, 2.0, 2.0, 1.0, x_arr, g, 70.0, 2, 0.1, 100000, 10000, 1e-6, 0)
  assert (logs.status == 0).all()
  fns_o = orc.set_up_example_fns(1, 1, 0)
  for b in range(B):
    p1, r1, a1, l1 = rx.solve_HJ_batch(1, 1, 1, epsl[b:b + 1], fns, nx, 1, nt, 2.0, 2.0, 1.0, x_arr, g[b:b + 1], 70.0, 2, 0.1, 100000, 10000,
                                       1e-6, 0)
    assert np.array_equal(p1[0], phi[b]) and np.array_equal(r1[0], rho[b]) and np.array_equal(l1.iters[0], logs.iters[b])
    info = {}
    res_o, _ = orc.solve_HJ(1, 1, 1, float(epsl[b]), fns_o, nx, 1, nt, 2.0, 2.0, 1.0, x_arr, 70.0, 2, 0.1, 100000, 10000, 1e-6, 0, g=g[b:b + 1],
                            info=info)
    assert info["block_iters"] == logs.iters[b].tolist()
    assert relmax(phi[b], res_o[0][1]) < TOL and relmax(rho[b], res_o[0][2]) < TOL and relmax(alp[b], res_o[0][3]) < TOL


def test_full_size_instance_properties(pk):
  """BASELINE configs[3] instance size (nx=1024, tsp=2), first 3 of 256 blocks: size-independent properties —
  phi[0] = g, rho >= 0, upwind masks (alp1 <= 0 <= alp2 since f = -a alp), and every converged block satisfies
  the implicit Engquist-Osher step of the HJ equation."""
  rx = pk["rx"]
  nx, nt_full = 1024, 257
  x_arr = rx.make_x_arr(1, nx, 1, 2.0, 2.0)
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, 1, 0)
  g = np.sin(np.pi * x_arr[0, :, 0])[None]
  T3 = 3.0 / (nt_full - 1)
  phi, rho, alp, logs = rx.solve_HJ_batch(1, 1, 1, [0.0], fns, nx, 1, 4, 2.0, 2.0, T3, x_arr, g, 70.0, 2, 0.1, 1000000, 10000, 1e-6, 0)
  assert logs.status[0] == 0 and (logs.end_reason[0] == 0).all()
  phi, rho, alp = phi[0], rho[0], alp[0]
  assert np.array_equal(phi[0], g[0]) and (rho >= 0).all()
  assert (alp[0] <= 0).all() and (alp[1] >= 0).all()
  dt, dx = 1.0 / (nt_full - 1), 2.0 / nx
  a = (x_arr[0, :, 0] - 1.0) ** 2 + 0.1
  for k in range(3):
    p1 = phi[k + 1]
    dp, dm = (np.roll(p1, -1) - p1) / dx, (p1 - np.roll(p1, 1)) / dx
    r = (p1 - phi[k]) / dt + a ** 2 / 2 * (np.minimum(dp, 0) ** 2 + np.maximum(dm, 0) ** 2)
    assert np.max(np.abs(r)) < 2e-2 * max(np.max(np.abs((p1 - phi[k]) / dt)), 1.0)


def test_device_tensors_stay_on_device(pk):
  import torch
  rx, upd = pk["rx"], pk["upd"]
  nx = 40
  x_arr = rx.make_x_arr(1, nx, 1, 2.0, 2.0)
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, 1, 0)
  d = golden("op_1d_eg1_K1")
  tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
  x24 = rx.make_x_arr(1, 24, 1, 2.0, 2.0)
  pn = upd.update_primal_1d(tt(d["phi"]), tt(d["rho"]), 70.0, tuple(tt(d["alp"][j]) for j in range(2)), float(d["tau"]), float(d["dt"]),
                            dspatial_of(d), fns, None, float(d["epsl"]), x24, None, 0)
  assert isinstance(pn, torch.Tensor) and pn.is_cuda
  assert relmax(pn.cpu().numpy(), d["phi_next"]) < TOL


def test_batched_2d_instances_cooperative_path(pk):
  """B = 3 independent 2-D instances (different epsl) through one handle of the cooperative kernel == separate solves."""
  rx = pk["rx"]
  nx, ny, nt = 16, 12, 3
  n_ctrl, bc, _ = rx.problem_setup(1, 2)
  x_arr = rx.make_x_arr(2, nx, ny, 2.0, 2.0)
  fns, _ = quiet(pk["sf"].set_up_example_fns, 1, 2, 0)
  g1 = pk["sf"].set_up_J(1, 2, (2.0, 2.0))(x_arr)
  g = np.concatenate([g1, 0.5 * g1, 1.5 * g1], axis=0)
  epsl = np.array([0.0, 0.05, 0.1])
  phi, rho, alp, logs = rx.solve_HJ_batch(2, n_ctrl, 1, epsl, fns, nx, ny, nt, 2.0, 2.0, 0.2, x_arr, g, 70.0, 2, 0.1, 400, 100, 1e-6, bc)
  assert phi.shape == (3, nt, nx, ny) and alp.shape == (3, 4, nt - 1, nx, ny, 2)
  for b in range(3):
    p1, r1, a1, l1 = rx.solve_HJ_batch(2, n_ctrl, 1, epsl[b:b + 1], fns, nx, ny, nt, 2.0, 2.0, 0.2, x_arr, g[b:b + 1], 70.0, 2, 0.1, 400, 100,
                                       1e-6, bc)
    assert np.array_equal(p1[0], phi[b]) and np.array_equal(a1[0], alp[b]) and np.array_equal(l1.iters[0], logs.iters[b])
  assert (alp[..., 1][:, :2] == 0).all() and (alp[..., 0][:, 2:] == 0).all()     # structurally-zero control components


def test_multi_step_with_device_tensors_and_save_middle(pk, tmp_path):
  import torch
  from pdhg_b200.solver import load_middle_solution
  rx, upd, sol = pk["rx"], pk["upd"], pk["sol"]
  d = golden("solve_1d_eg1_nx40_nt11")
  egno, ndim, nx, ny, n_ctrl, bc, x_arr, fns = _problem(pk, d)
  nt = int(d["nt"])
  g = torch.from_numpy(pk["sf"].set_up_J(1, 1, (2.0,))(x_arr)).cuda()
  (res, errs), _ = quiet(sol.PDHG_multi_step, upd.NativeUpdatePrimal(1, bc), upd.NativeUpdateDual(bc), fns, g, x_arr, 1, nt, (nx,), 1.0 / (nt - 1),
                         (2.0 / nx,), 70.0, time_step_per_PDHG=2, epsl=0.0, stepsz_param=0.1, n_ctrl=1, N_maxiter=1000000, print_freq=1000,
                         eps=1e-6, save_middle_dir=str(tmp_path), save_middle_prefix="mid")
  mi, phi, rho, alp = res[0]
  assert isinstance(phi, torch.Tensor) and phi.is_cuda
  assert relmax(phi.cpu().numpy(), d["phi"]) < TOL and relmax(alp.cpu().numpy(), d["alp"]) < TOL
  mid = load_middle_solution(str(tmp_path), "mid")       # [max_iters, phi_all, rho_all, alp_all, errs_all] (utils_pdhg_solver.py:211-212)
  assert mid[0] == int(d["max_iters"]) and len(mid[1]) == nt - 1 and len(mid[4]) == nt - 1


def test_run_example_cli_saves_reference_pickle_layout(pk, tmp_path):
  """python -m pdhg_b200.run_example with the reference's flags; the pickle holds (results, errs_all) (solver.py:13-26)."""
  import pickle, subprocess, sys, glob
  pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pdhg-optimal-control_b200")
  env = dict(os.environ, PYTHONPATH=pkg)
  r = subprocess.run([sys.executable, "-m", "pdhg_b200.run_example", "--ndim", "1", "--epsl", "0", "--egno", "1", "--nx", "40", "--nt", "11",
                      "--stepsz_param", "0.1", "--print_freq", "1000"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stderr[-2000:]
  assert "pdhg conv. Max err is" in r.stdout and "PDHG converges at iter 4615" in r.stdout
  files = glob.glob(str(tmp_path / "check_points" / "*" / "eg1_1d" / "nt11_nx40.pickle"))
  assert len(files) == 1
  results, errs_all = pickle.load(open(files[0], "rb"))
  d = golden("solve_1d_eg1_nx40_nt11")
  assert results[0][0] == int(d["max_iters"]) and relmax(results[0][1], d["phi"]) < TOL and len(errs_all) == 10

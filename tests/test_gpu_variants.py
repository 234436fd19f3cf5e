"""GPU: every variant of the cooperative kernel the benchmark times, forced onto grids the oracle can afford —
  * the TMA row pipelines (PDHG_TMA=1) against the reference goldens and, bit for bit, against the direct-load phases;
  * the fused dual sweeps (PDHG_DFUSE = 1, 2, 3, 5) on a cold start, where every outer iteration runs up to 10 inner sweeps and
    the exit test fires in the middle of a fused pass: bitwise-equal iterates and identical sweep counts across the settings;
  * the headline geometry itself (256 x 256, K = 64, stepsz 0.05; warp-private transforms, 16-deep Thomas, TMA) against the oracle.
Knobs are read once per handle (pdhg_create), so each setting gets its own handle (the cache key includes them)."""
import os

import numpy as np
import pytest

from helpers import TOL, golden, golden_names, quiet, relmax

pytestmark = pytest.mark.gpu

KNOBS = ("PDHG_FORCE_PATH", "PDHG_FORCE_W256", "PDHG_NO_W256", "PDHG_DFUSE", "PDHG_TMA", "PDHG_NO_K1", "PDHG_NO_BSLAB")


@pytest.fixture(scope="module")
def pk(built_lib):
  from pdhg_b200 import run_example, set_fns, update_fns_in_pdhg
  for k in KNOBS:
    os.environ.pop(k, None)
  yield dict(rx=run_example, sf=set_fns, upd=update_fns_in_pdhg)
  update_fns_in_pdhg.clear_handles()


class knobs:
  def __init__(self, **kw):
    self.kw = {k: str(v) for k, v in kw.items()}

  def __enter__(self):
    for k in KNOBS:
      os.environ.pop(k, None)
    os.environ.update(self.kw)

  def __exit__(self, *a):
    for k in KNOBS:
      os.environ.pop(k, None)


def _setup(pk, egno, ndim, nx, ny):
  rx = pk["rx"]
  n_ctrl, bc, cen = rx.problem_setup(egno, ndim)
  x_arr = rx.make_x_arr(ndim, nx, ny, 2.0, 2.0, cen)
  fns, _ = quiet(pk["sf"].set_up_example_fns, egno, ndim, 0)
  return n_ctrl, bc, x_arr, fns


def _solve(pk, egno, ndim, nx, ny, nt, T, tsp, epsl, stepsz, nmax, pf=10000, **kn):
  n_ctrl, bc, x_arr, fns = _setup(pk, egno, ndim, nx, ny)
  info = {}
  with knobs(PDHG_FORCE_PATH=2, **kn):
    (res, errs), log = quiet(pk["rx"].solve_HJ, ndim, n_ctrl, egno, epsl, fns, nx, ny, nt, 2.0, 2.0, T, x_arr, 70.0, tsp, stepsz, nmax, pf, 1e-6,
                             bc, info=info)
  return res, errs, info, log


SOLVES_2D = [n for n in golden_names("solve_2d")]


@pytest.mark.parametrize("name", SOLVES_2D)
def test_tma_pipelines_vs_reference_golden(pk, name):
  """2-D goldens (egno 1, 2, 3; periodic and Neumann-x; K = 1, 2; partial row tiles) with the TMA row pipelines forced on."""
  d = golden(name)
  egno, ndim, nx, ny, nt, tsp = [int(d[k]) for k in ("egno", "ndim", "nx", "ny", "nt", "tsp")]
  res, errs, info, _ = _solve(pk, egno, ndim, nx, ny, nt, 1.0, tsp, float(d["epsl"]), float(d["stepsz"]), int(d["N_maxiter"]),
                              int(d["print_freq"]), PDHG_TMA=1)
  mi, phi, rho, alp = res[0]
  assert info["block_iters"] == d["block_iters"].tolist() and info["stepsz_used"] == d["stepsz_used"].tolist()
  assert relmax(phi, d["phi"]) < TOL and relmax(rho, d["rho"]) < TOL and relmax(alp, d["alp"]) < TOL
  assert relmax(np.concatenate([e.reshape(-1, 2) for e in errs]), d["errs_flat"]) < 1e-7


@pytest.mark.parametrize("egno,nx,ny,nt,epsl,nmax", [(1, 30, 64, 4, 0.01, 60), (2, 17, 40, 3, 0.0, 50), (3, 22, 36, 2, 0.05, 40), (1, 64, 256, 5, 0.0, 30)])
def test_tma_pipelines_equal_direct_phases_bitwise(pk, egno, nx, ny, nt, epsl, nmax):
  """Same per-point arithmetic in both variants: the iterates must agree bit for bit (the error sums are reduced in a different
  order, so the logged errors agree to rounding only); odd nx leaves the last row tile partly empty."""
  tsp = nt if egno != 3 else 2
  T = (nt - 1) / 64.0
  a = _solve(pk, egno, 2, nx, ny, nt, T, tsp, epsl, 0.05, nmax, PDHG_TMA=1)
  b = _solve(pk, egno, 2, nx, ny, nt, T, tsp, epsl, 0.05, nmax, PDHG_TMA=0)
  assert a[2]["block_iters"] == b[2]["block_iters"] and a[2]["n_inner"] == b[2]["n_inner"]
  for x, y in zip(a[0][0][1:], b[0][0][1:]):
    assert np.array_equal(np.asarray(x), np.asarray(y))
  assert relmax(np.concatenate(a[1]), np.concatenate(b[1])) < 1e-9


@pytest.mark.parametrize("K,tma", [(3, 0), (3, 1), (16, 0), (16, 1)])
def test_fused_dual_sweeps_cold_start_bitwise_and_vs_oracle(pk, K, tma):
  """256 x 256, K = 3 and 16, warp-private transforms forced, cold start: the first outer iterations run 10 inner sweeps, later
  ones fewer (the exit test then fires in the middle of a fused pass and the pass is redone with the exact count).  The fused
  passes (2, 3, 5 sweeps per pass) must reproduce the sweep-by-sweep loop: bitwise-equal iterates, identical sweep totals."""
  from oracle import pdhg_numpy as orc
  nx = ny = 256
  nt, T = K + 1, K / 64.0
  nmax = 400 if K == 3 else 120
  runs = {}
  for df in ((1, 2, 3) if tma else (1, 2, 5)):
    runs[df] = _solve(pk, 1, 2, nx, ny, nt, T, nt, 0.0, 0.05, nmax, PDHG_FORCE_W256=1, PDHG_DFUSE=df, PDHG_TMA=tma)
  ref = runs[1]
  n_in = ref[2]["n_inner"]
  assert nmax < n_in < 10 * nmax, "the window must contain iterations with fewer than 10 and more than 1 inner sweeps"
  for df, r in runs.items():
    assert r[2]["n_inner"] == n_in and r[2]["block_iters"] == ref[2]["block_iters"], (df, r[2]["n_inner"], n_in)
    for x, y in zip(r[0][0][1:], ref[0][0][1:]):
      assert np.array_equal(np.asarray(x), np.asarray(y)), "DFUSE=%d differs from the sweep-by-sweep loop" % df
  # first iterations against the oracle (10 sweeps each)
  n0 = 4 if K == 3 else 2
  n_ctrl, bc, x_arr, fns = _setup(pk, 1, 2, nx, ny)
  res_o, errs_o = orc.solve_HJ(2, n_ctrl, 1, 0.0, orc.set_up_example_fns(1, 2, 0), nx, ny, nt, 2.0, 2.0, T, x_arr, 70.0, nt, 0.05, n0, 10000, 1e-6, bc)
  for df in runs:
    r = _solve(pk, 1, 2, nx, ny, nt, T, nt, 0.0, 0.05, n0, PDHG_FORCE_W256=1, PDHG_DFUSE=df, PDHG_TMA=tma)
    for x, y in zip(r[0][0][1:], res_o[0][1:]):
      assert relmax(x, y) < TOL
    assert relmax(r[1][0], errs_o[0]) < 1e-7


@pytest.mark.parametrize("K,ny", [(3, 256), (35, 64), (40, 32), (64, 16), (77, 8)])
def test_single_pass_phase_B_equals_three_pass_bitwise(pk, K, ny):
  """Phase B in one pass per ky-slab (x-FFT -> Thomas in shared memory -> inverse x-FFT; K = 40, 64, 77 need 2 or 3 k-chunks)
  runs the same butterflies and recurrences as the three grid-wide passes: bit-identical iterates."""
  nx, nt, T = 256, K + 1, K / 64.0
  a = _solve(pk, 1, 2, nx, ny, nt, T, nt, 0.01, 0.05, 12, PDHG_FORCE_W256=1)
  b = _solve(pk, 1, 2, nx, ny, nt, T, nt, 0.01, 0.05, 12, PDHG_FORCE_W256=1, PDHG_NO_BSLAB=1)
  assert a[2]["block_iters"] == b[2]["block_iters"] and a[2]["n_inner"] == b[2]["n_inner"]
  for x, y in zip(a[0][0][1:], b[0][0][1:]):
    assert np.array_equal(np.asarray(x), np.asarray(y))
  assert np.array_equal(np.concatenate(a[1]), np.concatenate(b[1]))


def test_headline_geometry_first_iterations_vs_oracle(pk):
  """The benchmarked configuration itself: BASELINE configs[2] with time_step_per_PDHG = 65 (256 x 256, K = 64, stepsz 0.05,
  default knobs: warp-private transforms, 16-deep Thomas, fused sweeps, TMA pipeline) — first 3 outer iterations against the
  oracle (about 90 s of NumPy)."""
  from oracle import pdhg_numpy as orc
  nx = ny = 256
  nt = 65
  n_ctrl, bc, x_arr, fns = _setup(pk, 1, 2, nx, ny)
  info = {}
  (res, errs), _ = quiet(pk["rx"].solve_HJ, 2, n_ctrl, 1, 0.0, fns, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, nt, 0.05, 3, 10000, 1e-6, bc, info=info)
  st = {}
  res_o, errs_o = orc.solve_HJ(2, n_ctrl, 1, 0.0, orc.set_up_example_fns(1, 2, 0), nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, nt, 0.05, 3, 10000, 1e-6, bc,
                               stats=st)
  assert info["block_iters"] == [3] and info["n_inner"] == st["n_inner"]
  for a, b in zip(res[0][1:], res_o[0][1:]):
    assert relmax(a, b) < TOL
  assert relmax(errs[0], errs_o[0]) < 1e-7


def test_baseline_cfg2_stepsz01_algorithm_failed_chain(pk):
  """BASELINE configs[1] (egno=1 ndim=1 epsl=0.1 nx=640 nt=161) at stepsz_param = 0.1: the reference NaNs at every step size of the
  fallback chain 0.1 -> 0.01 in block 0 and ends in 'algorithm failed' (utils_pdhg_solver.py:180-187; golden made from the
  reference sources by oracle/make_golden_baseline.py)."""
  d = golden("baseline_cfg2_stepsz01_failed")
  rx = pk["rx"]
  n_ctrl, bc, x_arr, fns = _setup(pk, 1, 1, 640, 1)
  for path in (1, 2):
    info = {}
    with knobs(PDHG_FORCE_PATH=path):
      (res, errs), log = quiet(rx.solve_HJ, 1, n_ctrl, 1, 0.1, fns, 640, 1, 161, 2.0, 2.0, 1.0, x_arr, 70.0, 2, 0.1, 1000000, 10000, 1e-6, bc, info=info)
    assert info["sol_nan"] and info["blocks_done"] == 0 and res[0][1] is None
    assert info["stepsz_final"] == d["stepsz_decrements"].tolist()[-1]
    announced = [float(l.rsplit(' ', 1)[1]) for l in log.splitlines() if 'decrease step size to' in l]
    assert announced == d["stepsz_decrements"].tolist()
    assert "algorithm failed" in log


def test_baseline_cfg3_blocks012_fallback_and_iterations(pk):
  """BASELINE configs[2] at the reference's default tsp = 2: time blocks 0..2 — block 1 NaNs at 0.1, 0.09, 0.08 and converges at
  0.07000000000000002; per-block iteration counts and solutions against the golden made from the reference sources."""
  gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
  # reference-generated fixture (oracle/make_golden_baseline.py cfg3, ~5 h through the shim) if present, else its oracle-generated twin
  name = next((n for n in ("baseline_cfg3_blocks012", "oracle_cfg3_blocks012") if os.path.exists(os.path.join(gdir, n + ".npz"))), None)
  if name is None:
    pytest.skip("fixture not generated")
  d = golden(name)
  nx = ny = 256
  nt, T = int(d["nt"]), float(d["T"])
  n_ctrl, bc, x_arr, fns = _setup(pk, 1, 2, nx, ny)
  info = {}
  (res, errs), log = quiet(pk["rx"].solve_HJ, 2, n_ctrl, 1, 0.0, fns, nx, ny, nt, 2.0, 2.0, T, x_arr, 70.0, 2, 0.1, 1000000, 10000, 1e-6, bc, info=info)
  mi, phi, rho, alp = res[0]
  assert info["stepsz_used"] == d["stepsz_used"].tolist()
  announced = [float(l.rsplit(' ', 1)[1]) for l in log.splitlines() if 'decrease step size to' in l]
  assert announced == d["stepsz_decrements"].tolist()
  s = int(d["sub"])
  assert relmax(phi, d["phi"]) < 1e-6          # (the stopping rule's own tolerance when the stopping iteration differs)
  it, it_ref = info["block_iters"], d["block_iters"].tolist()
  if it == it_ref:
    assert relmax(phi, d["phi"]) < TOL and relmax(np.asarray(rho)[:, ::s, ::s], d["rho_sub"]) < TOL
    assert relmax(np.asarray(alp)[:, :, ::s, ::s, :], d["alp_sub"]) < TOL
  assert it == it_ref, (it, it_ref)


@pytest.mark.parametrize("variant", ["k1", "cta", "coop"])
def test_error_log_overflow_does_not_stop_the_march(pk, variant):
  """max_rec = 3 with print_freq = 1: every block overflows its error log.  The march must still run all blocks (the status is
  PDHG_INST_LOG_OVERFLOW, only the extra rows are dropped) and produce the iterates of a run with a large log."""
  from pdhg_b200 import _lib
  from pdhg_b200.update_fns_in_pdhg import get_solver
  rx = pk["rx"]
  nx, nt = 40, 6
  n_ctrl, bc, x_arr, fns = _setup(pk, 1, 1, nx, 1)
  g = pk["sf"].set_up_J(1, 1, (2.0,))(x_arr)
  kn = dict(PDHG_FORCE_PATH=2) if variant == "coop" else (dict(PDHG_FORCE_PATH=1, PDHG_NO_K1=1) if variant == "cta" else dict(PDHG_FORCE_PATH=1))
  out = []
  for max_rec in (3, 4096):
    with knobs(**kn):
      s = get_solver(fns, (nx,), 1, bc, 1.0 / (nt - 1), (2.0 / nx,), 70.0, x_arr, batch=1, nblocks=nt - 1, max_rec=max_rec)
      out.append(s.multi_step_host(g, 0.0, 0.1, 3000, 1))
  (phi_a, rho_a, alp_a, la), (phi_b, rho_b, alp_b, lb) = out
  assert int(la.status[0]) == _lib.INST_LOG_OVERFLOW and int(lb.status[0]) == _lib.INST_OK
  assert int(la.blocks_done[0]) == nt - 1 == int(lb.blocks_done[0])
  assert la.iters.tolist() == lb.iters.tolist() and (la.nrec <= 3).all()
  assert np.array_equal(phi_a, phi_b) and np.array_equal(rho_a, rho_b) and np.array_equal(alp_a, alp_b)


@pytest.mark.parametrize("ndim,nx,ny,nt,tsp,epsl", [(1, 40, 1, 11, 2, 0.1), (2, 12, 12, 4, 2, 0.0), (1, 20, 1, 7, 3, 0.0)])
def test_resume_from_middle_file_equals_uninterrupted_march(pk, tmp_path, ndim, nx, ny, nt, tsp, epsl, monkeypatch):
  """--save_middle rewrites the middle file after EVERY time block (utils_pdhg_solver.py:211-212); a march restarted from the file
  an interrupted run left behind (after block 2 here; the viscous 1-D case has already fallen back to a smaller step size by
  then) must reproduce the uninterrupted march bit for bit: a working version of the reference's unwired load_middle path."""
  import shutil
  from pdhg_b200.utils import utils_pdhg_solver as sol
  from pdhg_b200 import solver as solver_mod
  n_ctrl, bc, x_arr, fns = _setup(pk, 1, ndim, nx, ny)
  nblocks = (nt - 1) // (tsp - 1)
  calls = []
  real_save = sol.save

  def spy_save(d, prefix, obj):
    real_save(d, prefix, obj)
    calls.append(len(obj[1]))
    if len(obj[1]) == 2:
      shutil.copy(os.path.join(d, prefix + ".pickle"), os.path.join(d, "interrupted.pickle"))
  monkeypatch.setattr(sol, "save", spy_save)
  info_a = {}
  (res_a, errs_a), _ = quiet(pk["rx"].solve_HJ, ndim, n_ctrl, 1, epsl, fns, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, tsp, 0.1, 1000000, 1000, 1e-6, bc,
                             save_middle_dir=str(tmp_path), save_middle_prefix="mid", info=info_a)
  assert calls == list(range(1, nblocks + 1))                     # one rewrite per block
  mid = solver_mod.load_middle_solution(str(tmp_path), "interrupted")
  assert len(mid[1]) == 2 and mid[5]["blocks_done"] == 2
  monkeypatch.setattr(sol, "save", real_save)
  info_b = {}
  (res_b, errs_b), log = quiet(pk["rx"].solve_HJ, ndim, n_ctrl, 1, epsl, fns, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, tsp, 0.1, 1000000, 1000, 1e-6, bc,
                               load_middle_dir=str(tmp_path), load_middle_prefix="interrupted", info=info_b)
  assert info_a["block_iters"] == info_b["block_iters"] and info_a["stepsz_used"] == info_b["stepsz_used"]
  for a, b in zip(res_a[0][1:], res_b[0][1:]):
    assert np.array_equal(np.asarray(a), np.asarray(b))
  assert len(errs_a) == len(errs_b) and all(np.array_equal(x, y) for x, y in zip(errs_a, errs_b))
  # the uninterrupted single-launch march gives the same answer as the block-wise one
  info_c = {}
  (res_c, _), _ = quiet(pk["rx"].solve_HJ, ndim, n_ctrl, 1, epsl, fns, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, tsp, 0.1, 1000000, 1000, 1e-6, bc, info=info_c)
  assert info_c["block_iters"] == info_a["block_iters"]
  for a, c in zip(res_a[0][1:], res_c[0][1:]):
    assert np.array_equal(np.asarray(a), np.asarray(c))

"""CPU: the NumPy oracle against the golden vectors produced from the reference sources (oracle/make_golden.py)."""
import numpy as np
import pytest

from helpers import TOL, dspatial_of, golden, golden_names, relmax
from oracle import pdhg_numpy as orc


@pytest.mark.parametrize("name", golden_names("op_"))
def test_operators_match_reference(name):
  d = golden(name)
  egno, ndim, nx, ny, K = [int(d[k]) for k in ("egno", "ndim", "nx", "ny", "K")]
  x_arr, bc, n_ctrl = orc.make_grid(egno, ndim, nx, ny, 2.0, 2.0)
  fns = orc.set_up_example_fns(egno, ndim, 0)
  dsp, dt = dspatial_of(d), float(d["dt"])
  nsp = (nx,) if ndim == 1 else (nx, ny)
  fv = orc.compute_Dxx_fft_fv(ndim, nsp, dsp, bc)
  assert relmax(fv.real, d["fv_real"]) < 1e-13
  alp = tuple(d["alp"][j] for j in range(2 * ndim))
  up = orc.update_primal_1d if ndim == 1 else orc.update_primal_2d
  with np.errstate(all="ignore"):
    res = orc.compute_cont_residual(d["rho"], alp, dt, dsp, fns, 70.0, float(d["epsl"]), x_arr, None, bc)
    pn = up(d["phi"], d["rho"], 70.0, alp, float(d["tau"]), dt, dsp, fns, fv, float(d["epsl"]), x_arr, None, bc,
            C=float(d["C"]), pow=float(d["pow"]), Ct=float(d["Ct"]))
    r1, a1, e1 = orc.update_dual_oneiter(d["phi_bar"], d["rho"], 70.0, alp, float(d["sigma"]), dt, dsp, float(d["epsl"]), x_arr,
                                         None, bc, fns, ndim)
    st = {}
    rN, aN = orc.update_dual_alternative(d["phi_bar"], d["rho"], 70.0, alp, float(d["sigma"]), dt, dsp, float(d["epsl"]), fns,
                                         x_arr, None, ndim, bc, eps=float(d["eps"]), stats=st)
  assert relmax(res, d["cont_residual"]) < 1e-13
  assert relmax(pn, d["phi_next"]) < 1e-13
  assert relmax(r1, d["rho_sweep1"]) < 1e-13 and relmax(np.stack(a1), d["alp_sweep1"]) < 1e-13
  eref = float(d["err_sweep1"])     # NaN for egno 3 (0/0 on the untouched y pair)
  assert (np.isnan(e1) and np.isnan(eref)) or abs(e1 - eref) <= 1e-13 * abs(eref)
  assert relmax(rN, d["rho_dual"]) < 1e-13 and relmax(np.stack(aN), d["alp_dual"]) < 1e-13
  assert st["n_inner"] == int(d["n_inner"])


FAST_SOLVES = ["solve_1d_eg1_nx20_nt6_tsp6", "solve_1d_eg1_nx24_nt9_tsp3_pow", "solve_2d_eg1_10x8_nt5_tsp3", "solve_2d_eg2_8x8_nt3",
               "solve_2d_eg3_10x12_nt3", "solve_1d_eg1_nx40_nt11_visc"]


@pytest.mark.parametrize("name", FAST_SOLVES)
def test_solve_matches_reference(name):
  d = golden(name)
  egno, ndim, nx, ny, nt, tsp = [int(d[k]) for k in ("egno", "ndim", "nx", "ny", "nt", "tsp")]
  x_arr, bc, n_ctrl = orc.make_grid(egno, ndim, nx, ny, 2.0, 2.0)
  fns = orc.set_up_example_fns(egno, ndim, 0)
  info = {}
  res, errs = orc.solve_HJ(ndim, n_ctrl, egno, float(d["epsl"]), fns, nx, ny, nt, 2.0, 2.0, 1.0, x_arr, 70.0, tsp, float(d["stepsz"]),
                           int(d["N_maxiter"]), int(d["print_freq"]), 1e-6, bc, C=float(d["C"]), pow=float(d["pow"]), Ct=float(d["Ct"]),
                           info=info)
  mi, phi, rho, alp = res[0]
  assert int(mi) == int(d["max_iters"])
  assert info["block_iters"] == d["block_iters"].tolist()
  assert info["stepsz_used"] == d["stepsz_used"].tolist()          # bit-exact fallback arithmetic
  assert relmax(phi, d["phi"]) < TOL and relmax(rho, d["rho"]) < TOL and relmax(alp, d["alp"]) < TOL
  assert [len(e) for e in errs] == d["errs_nrec"].tolist()
  assert relmax(np.concatenate([np.asarray(e).reshape(-1, 2) for e in errs]), d["errs_flat"]) < 1e-8


def test_algorithm_failed_path_matches_reference():
  d = golden("solve_1d_eg1_nx32_nt5_failed")
  x_arr, bc, n_ctrl = orc.make_grid(1, 1, 32, 1, 2.0, 2.0)
  info = {}
  res, errs = orc.solve_HJ(1, n_ctrl, 1, float(d["epsl"]), orc.set_up_example_fns(1, 1, 0), 32, 1, int(d["nt"]), 2.0, 2.0, 1.0, x_arr, 70.0, 2,
                           float(d["stepsz"]), int(d["N_maxiter"]), int(d["print_freq"]), 1e-6, bc, info=info)
  assert info["sol_nan"] and res[0][1] is None
  tried = [s for _, s in info["stepsz_tried"]]
  assert tried[1:] == d["stepsz_decrements"].tolist()


def test_oracle_run_of_cfg3_blocks_equals_reference_run():
  """BASELINE configs[2], time blocks 0-2 (11 104 iterations incl. three NaN fallbacks): the oracle's run
  (scripts/make_oracle_golden.py) against the run of the reference's own sources through the shim (oracle/make_golden_baseline.py,
  6.6 h of CPU).  Both fixtures are committed; every field they share must be identical, bit for bit."""
  import os
  gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
  ref = np.load(os.path.join(gdir, "baseline_cfg3_blocks012.npz"))
  orc = np.load(os.path.join(gdir, "oracle_cfg3_blocks012.npz"))
  assert ref["block_iters"].tolist() == [3810, 3704, 3590]                      # BASELINE.md / SURVEY App. C
  assert ref["stepsz_used"].tolist() == [0.1, 0.07000000000000002, 0.07000000000000002]
  for k in ("block_iters", "stepsz_used", "stepsz_decrements", "phi", "rho_sub", "alp_sub", "nt", "T", "sub"):
    assert np.array_equal(ref[k], orc[k]), k

"""Closed-loop trajectories (run_example.py:18-155): the NumPy restatement against goldens made from the reference's own
functions (CPU), and the CUDA kernel against both (GPU)."""
import numpy as np
import pytest

from helpers import golden, golden_names, quiet

TRAJ = golden_names("traj_")


@pytest.mark.parametrize("name", TRAJ)
def test_traj_oracle_vs_reference_golden(name):
  from oracle import pdhg_numpy as orc, traj_numpy as otr
  d = golden(name)
  egno, nt, epsl, method = int(d["egno"]), int(d["nt"]), float(d["epsl"]), str(d["method"])
  if int(d["ndim"]) == 1:
    ta, tx = otr.compute_traj_1d(d["x_init"], d["alp"], orc.set_up_example_fns(egno, 1, 0).f_fn, nt, d["x_arr"], d["t_arr"], float(d["x_period"]),
                                 float(d["T"]), epsl, method, noise=d["noise"])
  else:
    cen = bool(d["center"])
    ta, tx = otr.compute_traj_2d(d["x_init"], d["alp"], orc.set_up_example_fns(egno, 2, 0).f_fn, nt, d["x1_arr"], d["x2_arr"], d["t_arr"],
                                 float(d["x_period"]), float(d["y_period"]), float(d["T"]), tuple(int(b) for b in d["bc"]), (cen, cen), epsl, method,
                                 noise=d["noise"])
  assert np.max(np.abs(ta - d["traj_alp"])) < 1e-13 and np.max(np.abs(tx - d["traj_x"])) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("name", TRAJ)
def test_traj_kernel_vs_reference_golden(built_lib, name):
  from pdhg_b200 import run_example as rx, set_fns
  d = golden(name)
  egno, ndim, nt, epsl, method = int(d["egno"]), int(d["ndim"]), int(d["nt"]), float(d["epsl"]), str(d["method"])
  fns, _ = quiet(set_fns.set_up_example_fns, egno, ndim, 0)
  if ndim == 1:
    ta, tx = rx.compute_traj_1d(d["x_init"], d["alp"], fns, nt, d["x_arr"], d["t_arr"], float(d["x_period"]), float(d["T"]), epsl, method,
                                noise=d["noise"])
  else:
    cen = bool(d["center"])
    ta, tx = rx.compute_traj_2d(d["x_init"], d["alp"], fns, nt, d["x1_arr"], d["x2_arr"], d["t_arr"], float(d["x_period"]), float(d["y_period"]),
                                float(d["T"]), tuple(int(b) for b in d["bc"]), (cen, cen), epsl, method, noise=d["noise"])
  assert ta.shape == d["traj_alp"].shape and tx.shape == d["traj_x"].shape
  scale = max(1.0, float(np.max(np.abs(d["traj_x"]))))
  assert np.max(np.abs(tx - d["traj_x"])) < 1e-10 * scale and np.max(np.abs(ta - d["traj_alp"])) < 1e-10 * max(1.0, float(np.max(np.abs(d["traj_alp"]))))


@pytest.mark.gpu
def test_traj_batched_seeded_and_on_solver_output(built_lib):
  """A batch of 20 000 samples in one launch; the seeded draws are reproducible; trajectories under the control the solver
  returned for the README example stay bounded and follow the sign structure of the control (alp1 <= 0 <= alp2)."""
  from pdhg_b200 import run_example as rx, set_fns
  n_ctrl, bc, _ = rx.problem_setup(1, 1)
  nx, nt = 64, 9
  x_arr = rx.make_x_arr(1, nx, 1, 2.0, 2.0)
  fns, _ = quiet(set_fns.set_up_example_fns, 1, 1, 0)
  (res, _), _ = quiet(rx.solve_HJ, 1, n_ctrl, 1, 0.0, fns, nx, 1, nt, 2.0, 2.0, 1.0, x_arr, 70.0, 2, 0.1, 100000, 10000, 1e-6, bc)
  alp = np.asarray(res[0][3])[:, ::-1, :, 0]
  t_arr = np.linspace(0.0, 1.0, nt)
  x0 = np.linspace(0.0, 2.0, 20000)
  a1, x1 = rx.compute_traj_1d(x0, alp, fns, nt, x_arr[0, :, 0], t_arr, 2.0, 1.0, 0.01, seed=7)
  a2, x2 = rx.compute_traj_1d(x0, alp, fns, nt, x_arr[0, :, 0], t_arr, 2.0, 1.0, 0.01, seed=7)
  assert np.array_equal(x1, x2) and x1.shape == (nt, 20000) and a1.shape == (nt - 1, 20000, 1)
  assert np.all(np.isfinite(x1)) and np.max(np.abs(x1 - x0[None])) < 3.0

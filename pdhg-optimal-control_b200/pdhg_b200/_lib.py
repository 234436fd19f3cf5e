"""ctypes binding of the C ABI in include/pdhg_b200.h (libpdhg_b200.so, built in-tree by `build.py`).

There is no fallback: if the shared library is missing or no CUDA device is present every compute entry
raises.  PyTorch is only used by callers for device memory (`tensor.data_ptr()`); this module itself needs
nothing but ctypes + numpy.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PDHG_B200_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libpdhg_b200.so"))

PDHG_OK, PDHG_ERR_ARG, PDHG_ERR_CUDA, PDHG_ERR_NOMEM, PDHG_ERR_UNSUPPORTED = 0, -1, -2, -3, -4
INST_OK, INST_SOL_NAN, INST_PAUSED, INST_LOG_OVERFLOW = 0, 1, 3, 4
END_CONVERGED, END_NAN, END_MAXITER, END_PAUSED = 0, 1, 2, 3
LOG_COLS = 4

EXPORTS = ("pdhg_create", "pdhg_destroy", "pdhg_last_error", "pdhg_path", "pdhg_launch_count", "pdhg_last_kernel_ms", "pdhg_phase_times", "pdhg_debug_phase", "pdhg_ext_phase", "pdhg_update_primal",
           "pdhg_update_dual", "pdhg_solve_block", "pdhg_multi_step", "pdhg_multi_step_host", "pdhg_multi_step_range", "pdhg_get_march_state",
           "pdhg_set_march_state", "pdhg_compute_traj", "pdhg_max_fuse", "pdhg_ext_exchange_ok", "pdhg_ext_set_exchange")


class PdhgError(RuntimeError):
  def __init__(self, code, msg):
    super().__init__("pdhg_b200 error %d: %s" % (code, msg))
    self.code = code


class Config(C.Structure):
  _fields_ = [("ndim", C.c_int32), ("egno", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("K", C.c_int32),
              ("n_ctrl", C.c_int32), ("bc_x", C.c_int32), ("bc_y", C.c_int32),
              ("dt", C.c_double), ("dx", C.c_double), ("dy", C.c_double), ("c_on_rho", C.c_double),
              ("C", C.c_double), ("pow", C.c_double), ("Ct", C.c_double), ("eps", C.c_double),
              ("rho_alp_iters", C.c_int32), ("batch", C.c_int32), ("nblocks", C.c_int32), ("max_rec", C.c_int32),
              ("device", C.c_int32), ("path", C.c_int32)]


class ExtBuffers(C.Structure):
  _fields_ = [("phi_in", C.c_void_p), ("phi_out", C.c_void_p), ("phib", C.c_void_p), ("rho_in", C.c_void_p), ("alp_in", C.c_void_p),
              ("rho_out", C.c_void_p), ("alp_out", C.c_void_p), ("zt", C.c_void_p), ("sums", C.c_void_p)]


class Logs(C.Structure):
  _fields_ = [("iters", C.c_void_p), ("stepsz_used", C.c_void_p), ("nrec", C.c_void_p), ("errlog", C.c_void_p),
              ("end_reason", C.c_void_p), ("status", C.c_void_p), ("blocks_done", C.c_void_p),
              ("stepsz_final", C.c_void_p), ("inner_total", C.c_void_p)]


_lib = None


def load():
  """Loads libpdhg_b200.so (once).  Raises if it has not been built: there is no other implementation."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError("%s not found - run `python pdhg-optimal-control_b200/build.py` (no CPU fallback exists)" % LIB_PATH)
  lib = C.CDLL(LIB_PATH)
  vp, dp, i32, i64, dbl = C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_double
  lib.pdhg_create.restype = C.c_int
  lib.pdhg_create.argtypes = [C.POINTER(Config), dp, dp, C.POINTER(vp)]
  lib.pdhg_destroy.restype = None
  lib.pdhg_destroy.argtypes = [vp]
  lib.pdhg_last_error.restype = C.c_char_p
  lib.pdhg_last_error.argtypes = []
  lib.pdhg_path.restype = C.c_int
  lib.pdhg_path.argtypes = [vp]
  lib.pdhg_last_kernel_ms.restype = dbl
  lib.pdhg_last_kernel_ms.argtypes = [vp]
  lib.pdhg_phase_times.restype = C.c_int
  lib.pdhg_phase_times.argtypes = [vp, dp]
  lib.pdhg_debug_phase.restype = C.c_int
  lib.pdhg_debug_phase.argtypes = [vp, C.c_int, C.c_int, dbl, C.c_int]
  lib.pdhg_ext_phase.restype = C.c_int
  lib.pdhg_ext_phase.argtypes = [vp, C.c_int, C.c_int, dbl, dbl, C.POINTER(ExtBuffers), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]
  lib.pdhg_launch_count.restype = i64
  lib.pdhg_launch_count.argtypes = [vp]
  lib.pdhg_update_primal.restype = C.c_int
  lib.pdhg_update_primal.argtypes = [vp, dp, dp, dp, dp, dbl, dp, vp]
  lib.pdhg_update_dual.restype = C.c_int
  lib.pdhg_update_dual.argtypes = [vp, dp, dp, dp, dp, dbl, dbl, dp, dp, dp, dp, vp]
  lib.pdhg_solve_block.restype = C.c_int
  lib.pdhg_solve_block.argtypes = [vp, dp, dp, dp, dp, dp, i64, i64, i64, i32, dp, dp, dp, C.POINTER(Logs), vp]
  lib.pdhg_multi_step.restype = C.c_int
  lib.pdhg_multi_step.argtypes = [vp, dp, dp, dp, i64, i32, dp, dp, dp, C.POINTER(Logs), vp]
  lib.pdhg_multi_step_range.restype = C.c_int
  lib.pdhg_multi_step_range.argtypes = [vp, dp, dp, dp, dp, i64, i32, i32, i32, dp, dp, dp, C.POINTER(Logs), vp]
  lib.pdhg_get_march_state.restype = C.c_int
  lib.pdhg_get_march_state.argtypes = [vp, dp, dp, dp, vp]
  lib.pdhg_set_march_state.restype = C.c_int
  lib.pdhg_set_march_state.argtypes = [vp, dp, dp, dp, vp]
  lib.pdhg_max_fuse.restype = C.c_int
  lib.pdhg_max_fuse.argtypes = [vp]
  lib.pdhg_ext_exchange_ok.restype = C.c_int
  lib.pdhg_ext_exchange_ok.argtypes = [vp]
  lib.pdhg_ext_set_exchange.restype = C.c_int
  lib.pdhg_ext_set_exchange.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
  lib.pdhg_compute_traj.restype = C.c_int
  lib.pdhg_compute_traj.argtypes = [i32] * 10 + [dbl, dbl, dbl] + [dp] * 8 + [vp]
  lib.pdhg_multi_step_host.restype = C.c_int
  lib.pdhg_multi_step_host.argtypes = [vp, dp, dp, dp, i64, i32, dp, dp, dp, C.POINTER(Logs)]
  _lib = lib
  return lib


def _check(rc):
  if rc != 0:
    raise PdhgError(rc, load().pdhg_last_error().decode("utf-8", "replace"))


def _hptr(a):
  return a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
  a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
  if shape is not None:
    a = np.ascontiguousarray(np.broadcast_to(a, shape))
  return a


class LogBuffers:
  """Host arrays behind a `pdhg_logs` struct."""

  def __init__(self, B, nblocks, max_rec):
    self.iters = np.zeros((B, nblocks), np.int64)
    self.stepsz_used = np.zeros((B, nblocks), np.float64)
    self.nrec = np.zeros((B, nblocks), np.int32)
    self.errlog = np.zeros((B, nblocks, max_rec, LOG_COLS), np.float64)
    self.end_reason = np.zeros((B, nblocks), np.int32)
    self.status = np.zeros((B,), np.int32)
    self.blocks_done = np.zeros((B,), np.int32)
    self.stepsz_final = np.zeros((B,), np.float64)
    self.inner_total = np.zeros((B,), np.int64)
    self.struct = Logs(*[_hptr(getattr(self, f[0])) for f in Logs._fields_])


def compute_traj_dev(ndim, egno, n_ctrl, nx, ny, nt, n_sample, bc_x, bc_y, nearest, x_period, y_period, epsl, alp_ptr, xn_ptr, yn_ptr, t_ptr,
                     noise_ptr, x0_ptr, traj_x_ptr, traj_alp_ptr, stream=None):
  """pdhg_compute_traj on device pointers (ints from `tensor.data_ptr()`)."""
  rc = load().pdhg_compute_traj(int(ndim), int(egno), int(n_ctrl), int(nx), int(ny), int(nt), int(n_sample), int(bc_x), int(bc_y), int(nearest),
                                float(x_period), float(y_period), float(epsl), alp_ptr, xn_ptr, yn_ptr, t_ptr, noise_ptr, x0_ptr, traj_x_ptr,
                                traj_alp_ptr, stream)
  if rc != 0:
    raise PdhgError(rc, "pdhg_compute_traj: bad argument or launch failure")


class Solver:
  """One `pdhg_handle`: a problem family (grid, egno, preconditioner) x a batch of B independent instances."""

  def __init__(self, ndim, egno, nx, ny, K, n_ctrl, bc, dt, dx, dy, c_on_rho, coef_x, coef_y=None, C_=1.0, pow_=1.0,
               Ct=1.0, eps=1e-6, rho_alp_iters=10, batch=1, nblocks=1, max_rec=128, device=0, path=0):
    lib = load()
    bc_x, bc_y = (bc, 0) if ndim == 1 else bc
    self.cfg = Config(ndim, egno, nx, ny if ndim == 2 else 1, K, n_ctrl, bc_x, bc_y, dt, dx, dy, c_on_rho, C_, pow_, Ct,
                      eps, rho_alp_iters, batch, nblocks, max_rec, device, path)
    self._coef_x = _f64(coef_x).ravel()
    self._coef_y = None if coef_y is None else _f64(coef_y).ravel()
    assert self._coef_x.size == nx and (self._coef_y is None or self._coef_y.size == ny)
    h = C.c_void_p()
    _check(lib.pdhg_create(C.byref(self.cfg), _hptr(self._coef_x), None if self._coef_y is None else _hptr(self._coef_y),
                           C.byref(h)))
    self._h = h
    self.lib = lib
    self.B, self.K, self.nblocks, self.max_rec = batch, K, nblocks, max_rec
    self.n = nx * (ny if ndim == 2 else 1)
    self.spatial = (nx,) if ndim == 1 else (nx, ny)
    self.A, self.n_ctrl, self.ndim = 2 * ndim, n_ctrl, ndim
    self.nt = nblocks * K + 1

  def close(self):
    if getattr(self, "_h", None):
      self.lib.pdhg_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  @property
  def path(self):
    return self.lib.pdhg_path(self._h)

  @property
  def last_kernel_ms(self):
    return float(self.lib.pdhg_last_kernel_ms(self._h))

  def phase_times_ms(self):
    out = np.zeros(16)
    _check(self.lib.pdhg_phase_times(self._h, _hptr(out)))
    names = ("A_residual_ffty", "B_fftx_tsolve", "C_iffty_phi", "D_dual_reduce", "unused", "setup_records_output",
             "a_compute", "a_fft", "a_store", "b_pass1", "b_pass2", "d_compute", "c_load", "c_fft", "c_update", "d_reduce")
    return dict(zip(names, (out / 1e6).tolist()))

  def ext_phase(self, phase, step, epsl, sum_lo, sum_hi, stream=None, pass_mask=7, nyh_override=0, ky_off=0, nyh_tab=0, **bufs):
    """One phase of the cooperative kernel on caller-owned device buffers (ints from `tensor.data_ptr()`); slab mode."""
    eb = ExtBuffers(*[bufs.get(f[0]) for f in ExtBuffers._fields_])
    _check(self.lib.pdhg_ext_phase(self._h, int(phase), int(pass_mask), float(step), float(epsl), C.byref(eb), int(sum_lo), int(sum_hi),
                                   int(nyh_override), int(ky_off), int(nyh_tab), stream))

  @property
  def exchange_ok(self):
    return bool(self.lib.pdhg_ext_exchange_ok(self._h))

  def set_exchange(self, P, rank, nxl, kyl, nyh, ptrs, gather=False):
    """Fused transposes of the slab mode (pdhg_ext_set_exchange): `ptrs[d]` = device pointer (int) of rank d's buffer as seen from this
    device; `gather`: phase 1 also loads its input rows from there; P = 0 switches the exchange off."""
    arr = (C.c_void_p * 8)(*([int(p) for p in ptrs] + [None] * (8 - len(ptrs)))) if P else None
    _check(self.lib.pdhg_ext_set_exchange(self._h, int(P), int(rank), int(nxl), int(kyl), int(nyh), int(bool(gather)), arr))

  @property
  def max_fuse(self):
    return int(self.lib.pdhg_max_fuse(self._h))

  def debug_phase(self, phase, pass_mask=7, step=0.05, reps=1):
    _check(self.lib.pdhg_debug_phase(self._h, int(phase), int(pass_mask), float(step), int(reps)))

  @property
  def launch_count(self):
    return int(self.lib.pdhg_launch_count(self._h))

  # ---- host-buffer march (PDHG_multi_step for B instances) ----
  def multi_step_host(self, g, epsl, stepsz, n_maxiter, print_freq):
    B = self.B
    g = _f64(g, (B,) + self.spatial)
    epsl = _f64(epsl, (B,))
    stepsz = _f64(stepsz, (B,))
    phi = np.empty((B, self.nt) + self.spatial)
    rho = np.empty((B, self.nt - 1) + self.spatial)
    alp = np.empty((B, self.A, self.nt - 1) + self.spatial + (self.n_ctrl,))
    logs = LogBuffers(B, self.nblocks, self.max_rec)
    _check(self.lib.pdhg_multi_step_host(self._h, _hptr(g), _hptr(epsl), _hptr(stepsz), int(n_maxiter), int(print_freq),
                                         _hptr(phi), _hptr(rho), _hptr(alp), C.byref(logs.struct)))
    return phi, rho, alp, logs

  # ---- device-buffer entry points (pointers are ints from torch `data_ptr()`) ----
  def multi_step_dev(self, g_ptr, epsl, stepsz, n_maxiter, print_freq, phi_ptr, rho_ptr, alp_ptr, stream=None):
    logs = LogBuffers(self.B, self.nblocks, self.max_rec)
    epsl = _f64(epsl, (self.B,))
    stepsz = _f64(stepsz, (self.B,))
    _check(self.lib.pdhg_multi_step(self._h, g_ptr, _hptr(epsl), _hptr(stepsz), int(n_maxiter), int(print_freq),
                                    phi_ptr, rho_ptr, alp_ptr, C.byref(logs.struct), stream))
    return logs

  def multi_step_range_dev(self, g_ptr, epsl, stepsz0, stepsz_cur, n_maxiter, print_freq, blk_begin, blk_end, phi_ptr, rho_ptr, alp_ptr,
                           logs=None, stream=None):
    """Time blocks [blk_begin, blk_end) of the march (pdhg_multi_step_range); `logs` may be passed back in to accumulate."""
    if logs is None:
      logs = LogBuffers(self.B, self.nblocks, self.max_rec)
    epsl = _f64(epsl, (self.B,))
    stepsz0 = _f64(stepsz0, (self.B,))
    cur = None if stepsz_cur is None else _f64(stepsz_cur, (self.B,))
    _check(self.lib.pdhg_multi_step_range(self._h, g_ptr, _hptr(epsl), _hptr(stepsz0), None if cur is None else _hptr(cur), int(n_maxiter),
                                          int(print_freq), int(blk_begin), int(blk_end), phi_ptr, rho_ptr, alp_ptr, C.byref(logs.struct),
                                          stream))
    return logs

  def get_march_state(self, phi_ptr, rho_ptr, alp_ptr, stream=None):
    _check(self.lib.pdhg_get_march_state(self._h, phi_ptr, rho_ptr, alp_ptr, stream))

  def set_march_state(self, phi_ptr, rho_ptr, alp_ptr, stream=None):
    _check(self.lib.pdhg_set_march_state(self._h, phi_ptr, rho_ptr, alp_ptr, stream))

  def solve_block_dev(self, phi0_ptr, rho0_ptr, alp0_ptr, epsl, stepsz, n_maxiter, iter_begin, iter_pause, print_freq,
                      phi_out_ptr, rho_out_ptr, alp_out_ptr, stream=None):
    logs = LogBuffers(self.B, 1, self.max_rec)
    epsl = _f64(epsl, (self.B,))
    stepsz = _f64(stepsz, (self.B,))
    _check(self.lib.pdhg_solve_block(self._h, phi0_ptr, rho0_ptr, alp0_ptr, _hptr(epsl), _hptr(stepsz), int(n_maxiter),
                                     int(iter_begin), int(iter_pause), int(print_freq), phi_out_ptr, rho_out_ptr,
                                     alp_out_ptr, C.byref(logs.struct), stream))
    return logs

  def update_primal_dev(self, phi_prev_ptr, rho_prev_ptr, alp_prev_ptr, epsl, tau, phi_next_ptr, stream=None):
    epsl = _f64(epsl, (self.B,))
    _check(self.lib.pdhg_update_primal(self._h, phi_prev_ptr, rho_prev_ptr, alp_prev_ptr, _hptr(epsl), float(tau),
                                       phi_next_ptr, stream))

  def update_dual_dev(self, phi_bar_ptr, rho_prev_ptr, alp_prev_ptr, epsl, sigma, eps, rho_next_ptr, alp_next_ptr,
                      stream=None):
    epsl = _f64(epsl, (self.B,))
    n_inner = np.zeros((self.B,), np.int32)
    err = np.zeros((self.B,), np.float64)
    _check(self.lib.pdhg_update_dual(self._h, phi_bar_ptr, rho_prev_ptr, alp_prev_ptr, _hptr(epsl), float(sigma),
                                     float(eps), rho_next_ptr, alp_next_ptr, _hptr(n_inner), _hptr(err), stream))
    return n_inner, err

"""Update operators (mirror of jaxsrc/update_fns_in_pdhg.py) backed by the CUDA kernels.

`update_primal_1d/2d`, `update_dual_oneiter`, `update_dual_alternative` keep the reference's names, argument
order and meaning.  Arrays may be NumPy (copied to the GPU and back) or torch CUDA tensors (stay on device).
`NativeUpdatePrimal` / `NativeUpdateDual` are the callables `run_example.solve_HJ` builds in place of the
reference's two lambdas (run_example.py:193-203); `PDHG_solver_oneiter` / `PDHG_multi_step` recognise them and
run the whole loop on the GPU instead of calling them once per iteration.
"""
import collections
import os

import numpy as np

from . import _dev, _lib
from .set_fns import coef_tables

# Handle cache: an LRU of at most PDHG_HANDLE_CACHE (default 8) `pdhg_handle`s — a handle pins its workspaces in device memory
# (0.7 GB for the 256 x 256 x 64 space-time block), so sweeps over grids / epsl / batch sizes must not accumulate them.
_handles = collections.OrderedDict()
_KNOB_ENV = ("PDHG_NO_W256", "PDHG_FORCE_W256", "PDHG_DFUSE", "PDHG_TMA", "PDHG_NO_BSLAB", "PDHG_NO_K1", "PDHG_PROFILE", "PDHG_MAX_RADIX")


def _knob_key():
  """The library reads its diagnostic knobs from the environment once per handle (pdhg_create), so they are part of the key."""
  return tuple(os.environ.get(k) for k in _KNOB_ENV)


def get_solver(fns_dict, nspatial, K, bc, dt, dspatial, c_on_rho, x_arr, C=1.0, pow=1.0, Ct=1.0, eps=1e-6, rho_alp_iters=10,
               batch=1, nblocks=1, max_rec=128, device=None, path=0):
  """Cached `pdhg_handle` for a (problem, grid, preconditioner) family on `device` (default: the current CUDA device)."""
  if device is None:
    device = _dev.current_device()
  ndim = fns_dict.ndim
  path = int(os.environ.get("PDHG_FORCE_PATH", path))   # testing hook: 1 single-CTA kernel, 2 cooperative kernel
  nx = int(nspatial[0])
  ny = int(nspatial[1]) if ndim == 2 else 1
  dx = float(dspatial[0])
  dy = float(dspatial[1]) if ndim == 2 else 1.0
  coef_x, coef_y = coef_tables(fns_dict.egno, ndim, _np(x_arr))
  key = (ndim, fns_dict.egno, nx, ny, int(K), fns_dict.n_ctrl, bc if ndim == 1 else tuple(bc), float(dt), dx, dy, float(c_on_rho),
         float(C), float(pow), float(Ct), float(eps), int(rho_alp_iters), int(batch), int(nblocks), int(max_rec), device, path,
         coef_x.tobytes(), None if coef_y is None else coef_y.tobytes(), _knob_key())
  s = _handles.get(key)
  if s is not None:
    _handles.move_to_end(key)
  else:
    s = _lib.Solver(ndim, fns_dict.egno, nx, ny, int(K), fns_dict.n_ctrl, bc, float(dt), dx, dy, float(c_on_rho), coef_x, coef_y,
                    float(C), float(pow), float(Ct), float(eps), int(rho_alp_iters), int(batch), int(nblocks), int(max_rec),
                    device, path)
    _handles[key] = s
    cap = max(1, int(os.environ.get("PDHG_HANDLE_CACHE", "8")))
    while len(_handles) > cap:
      _handles.popitem(last=False)      # least recently used; destroyed (Solver.__del__) once no caller holds it any more
  return s


def clear_handles():
  for s in _handles.values():
    s.close()
  _handles.clear()


def _np(x):
  return x.detach().cpu().numpy() if _dev.is_tensor(x) else np.asarray(x)


def _stack_alp(alp):
  """tuple of [K,...,n_ctrl] -> one device tensor [A,K,...,n_ctrl] (the C ABI / returned-array layout)."""
  t = _dev.require_cuda()
  return t.stack([_dev.to_dev(a) for a in alp], dim=0).contiguous()


def _update_primal(ndim, phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc, C, pow, Ct):
  t = _dev.require_cuda()
  K = rho_prev.shape[0]
  s = get_solver(fns_dict, rho_prev.shape[1:], K, bc, dt, dspatial, c_on_rho, x_arr, C=C, pow=pow, Ct=Ct)
  phi_d, rho_d, alp_d = _dev.to_dev(phi_prev), _dev.to_dev(rho_prev), _stack_alp(alp_prev)
  out = t.empty_like(phi_d)
  s.update_primal_dev(phi_d.data_ptr(), rho_d.data_ptr(), alp_d.data_ptr(), float(epsl), float(tau), out.data_ptr(),
                      _dev.stream_ptr())
  return _dev.like_input(out, phi_prev)


def update_primal_1d(phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc,
                     C=1.0, pow=1, Ct=1):
  """phi_next = phi_prev + tau * H1_precond_1d(cont_residual)   (update_fns_in_pdhg.py:135-140)."""
  return _update_primal(1, phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc, C, pow, Ct)


def update_primal_2d(phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc,
                     C=1.0, pow=1, Ct=1):
  """2-D primal step; `pow` and `Ct` are ignored exactly as in the reference (update_fns_in_pdhg.py:142-147)."""
  return _update_primal(2, phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, bc, C, 1.0, 1.0)


def _update_dual(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr, ndim, bc, rho_alp_iters, eps):
  t = _dev.require_cuda()
  K = rho_prev.shape[0]
  s = get_solver(fns_dict, rho_prev.shape[1:], K, bc, dt, dspatial, c_on_rho, x_arr, rho_alp_iters=rho_alp_iters)
  phi_d, rho_d, alp_d = _dev.to_dev(phi_bar), _dev.to_dev(rho_prev), _stack_alp(alp_prev)
  rho_o, alp_o = t.empty_like(rho_d), t.empty_like(alp_d)
  n_inner, err = s.update_dual_dev(phi_d.data_ptr(), rho_d.data_ptr(), alp_d.data_ptr(), float(epsl), float(sigma), float(eps),
                                   rho_o.data_ptr(), alp_o.data_ptr(), _dev.stream_ptr())
  rho_next = _dev.like_input(rho_o, rho_prev)
  alp_next = tuple(_dev.like_input(alp_o[j], alp_prev[0]) for j in range(alp_o.shape[0]))
  return rho_next, alp_next, float(err[0]), int(n_inner[0])


def update_dual_oneiter(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, x_arr, t_arr, bc, fns_dict, ndim):
  """One alp->rho sweep and its squared relative change `err` (update_fns_in_pdhg.py:150-165)."""
  rho_next, alp_next, err, _ = _update_dual(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr,
                                            ndim, bc, 1, -1.0)
  return rho_next, alp_next, err


def update_dual_alternative(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr, ndim, bc,
                            rho_alp_iters=10, eps=1e-7):
  """Up to `rho_alp_iters` sweeps with the prox centre moving, early exit on err < eps (update_fns_in_pdhg.py:167-180)."""
  rho_next, alp_next, _, _ = _update_dual(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr,
                                          ndim, bc, rho_alp_iters, eps)
  return rho_next, alp_next


class NativeUpdatePrimal:
  """Stands in for the `fn_update_primal` lambda of run_example.py:193-195,199-201 (same call signature)."""

  def __init__(self, ndim, bc, C=1.0, pow=1.0, Ct=1.0):
    self.ndim, self.bc, self.C, self.pow, self.Ct = ndim, bc, float(C), float(pow), float(Ct)
    if ndim == 2:   # update_fns_in_pdhg.py:146: "pow and Ct are not implemented"
      self.pow, self.Ct = 1.0, 1.0

  def __call__(self, phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr):
    fn = update_primal_1d if self.ndim == 1 else update_primal_2d
    return fn(phi_prev, rho_prev, c_on_rho, alp_prev, tau, dt, dspatial, fns_dict, fv, epsl, x_arr, t_arr, self.bc,
              C=self.C, pow=self.pow, Ct=self.Ct)


class NativeUpdateDual:
  """Stands in for the `fn_update_dual` lambda of run_example.py:196-197,202-203."""

  def __init__(self, bc, rho_alp_iters=10):
    self.bc, self.rho_alp_iters = bc, rho_alp_iters

  def __call__(self, phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr, ndim, eps=1e-7):
    return update_dual_alternative(phi_bar, rho_prev, c_on_rho, alp_prev, sigma, dt, dspatial, epsl, fns_dict, x_arr, t_arr, ndim,
                                   self.bc, rho_alp_iters=self.rho_alp_iters, eps=eps)

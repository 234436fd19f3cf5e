"""Host-side setup mirror of jaxsrc/utils/utils_precond.py:42-71.

`compute_Dxx_fft_fv` is evaluated once per solve in the reference (run_example.py:191) and handed to the update
operators as `fv`.  The CUDA library builds the same symbol in closed form at `pdhg_create` time
((2cos(2 pi k/n)-2)/h^2), so `fv` is accepted by the mirrored signatures but not consumed.
This function is kept for API compatibility and returns that closed form."""
import numpy as np


def compute_Dxx_fft_fv(ndim, nspatial, dspatial, bc):
  if ndim == 1:
    if bc != 0:
      raise NotImplementedError
    k = np.arange(nspatial[0])
    return ((2 * np.cos(2 * np.pi * k / nspatial[0]) - 2) / dspatial[0] ** 2).astype(np.complex128)
  if ndim == 2:
    if tuple(bc) != (0, 0):
      raise NotImplementedError
    kx = np.arange(nspatial[0])[:, None]
    ky = np.arange(nspatial[1])[None, :]
    return ((2 * np.cos(2 * np.pi * kx / nspatial[0]) - 2) / dspatial[0] ** 2
            + (2 * np.cos(2 * np.pi * ky / nspatial[1]) - 2) / dspatial[1] ** 2).astype(np.complex128)
  raise NotImplementedError

"""Shared helpers for the parity tests."""
import contextlib
import io
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp64 parity bar of BASELINE.json's north_star: relative L-inf <= 1e-10 on phi, rho, alp
TOL = 1e-10


def relmax(a, b):
  a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
  assert a.shape == b.shape, (a.shape, b.shape)
  if a.size == 0:
    return 0.0
  assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
  m = ~np.isnan(b)
  if not m.any():
    return 0.0
  return float(np.max(np.abs(a[m] - b[m])) / max(np.max(np.abs(b[m])), 1e-300))


def quiet(fn, *a, **k):
  buf = io.StringIO()
  with contextlib.redirect_stdout(buf):
    out = fn(*a, **k)
  return out, buf.getvalue()


def golden(name):
  return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_names(prefix):
  return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def dspatial_of(d):
  nx, ny, ndim = int(d["nx"]), int(d["ny"]), int(d["ndim"])
  return (2.0 / nx,) if ndim == 1 else (2.0 / nx, 2.0 / ny)

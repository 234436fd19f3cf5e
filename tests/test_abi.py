"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/pdhg_b200.h declares.
No compute is called (no GPU here) except to check that it FAILS loudly without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  hdr = open(os.path.join(ROOT, "include", "pdhg_b200.h")).read()
  hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
  return sorted(set(re.findall(r"\b(pdhg_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(built_lib):
  lib = ctypes.CDLL(built_lib)
  names = _declared_symbols()
  assert len(names) >= 10
  for n in names:
    assert hasattr(lib, n), "missing export " + n


def test_binding_covers_header(built_lib):
  from pdhg_b200 import _lib
  assert sorted(_lib.EXPORTS) == _declared_symbols()
  lib = _lib.load()
  for n in _lib.EXPORTS:
    assert getattr(lib, n).restype is not ctypes.c_int or getattr(lib, n).argtypes is not None


def test_config_struct_layout(built_lib):
  from pdhg_b200 import _lib
  # 8 int32, 8 doubles, 6 int32 -> 32 + 64 + 24 = 120 bytes, natural alignment, no padding surprises
  assert ctypes.sizeof(_lib.Config) == 120
  assert _lib.Config.dt.offset == 32 and _lib.Config.rho_alp_iters.offset == 96
  assert ctypes.sizeof(_lib.Logs) == 9 * ctypes.sizeof(ctypes.c_void_p)


def test_sass_is_sm100a(built_lib):
  import subprocess
  out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
  assert "sm_100a" in out


def test_no_cpu_fallback(built_lib):
  """Without a CUDA device creation must fail with PDHG_ERR_CUDA — the product never computes on the CPU."""
  import torch
  if torch.cuda.is_available():
    pytest.skip("a GPU is present")
  from pdhg_b200 import _lib
  with pytest.raises(_lib.PdhgError) as ei:
    _lib.Solver(1, 1, 16, 1, 1, 1, 0, 0.1, 0.125, 1.0, 70.0, np.ones(16))
  assert ei.value.code == _lib.PDHG_ERR_CUDA


def test_bad_arguments_are_rejected(built_lib):
  from pdhg_b200 import _lib
  lib = _lib.load()
  h = ctypes.c_void_p()
  assert lib.pdhg_create(None, None, None, ctypes.byref(h)) == _lib.PDHG_ERR_ARG
  cfg = _lib.Config(3, 1, 16, 1, 1, 1, 0, 0, 0.1, 0.1, 1.0, 70.0, 1.0, 1.0, 1.0, 1e-6, 10, 1, 1, 8, 0, 0)
  c = np.ones(16)
  assert lib.pdhg_create(ctypes.byref(cfg), c.ctypes.data_as(ctypes.c_void_p), None, ctypes.byref(h)) == _lib.PDHG_ERR_ARG
  assert b"ndim" in lib.pdhg_last_error()


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, "pdhg-optimal-control_b200")
  for dp, _, fs in os.walk(pkg):
    for f in fs:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        src = open(os.path.join(dp, f)).read()
        assert "oracle" not in src.replace("no oracle", ""), os.path.join(dp, f)

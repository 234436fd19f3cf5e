import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pdhg-optimal-control_b200")
for p in (ROOT, PKG):
  if p not in sys.path:
    sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
  """Path of libpdhg_b200.so, (re)built in-tree if stale.  nvcc cross-compiles without a GPU."""
  import importlib.util
  spec = importlib.util.spec_from_file_location("pdhg_b200_build", os.path.join(PKG, "build.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod.build()


def pytest_collection_modifyitems(config, items):
  # GPU tests are selected explicitly with -m gpu; skip them when no device is visible.
  try:
    import torch
    has_gpu = torch.cuda.is_available()
  except Exception:
    has_gpu = False
  if has_gpu:
    return
  skip = pytest.mark.skip(reason="no CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)

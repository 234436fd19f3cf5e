"""GPU: the raw ctypes stub printed in INTEGRATION.md section 3 must run as written (only the library path is patched)."""
import os
import re

import numpy as np
import pytest

from helpers import golden, relmax

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_integration_md_ctypes_stub_runs(built_lib):
  md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
  sec = md.split("## 3. Raw C ABI")[1]
  code = re.search(r"```python\n(.*?)```", sec, re.S).group(1)
  code = code.replace('".../pdhg-optimal-control_b200/lib/libpdhg_b200.so"', repr(built_lib))
  ns = {}
  exec(compile(code, "INTEGRATION.md#3", "exec"), ns)
  d = golden("oracle_cfg1")
  assert ns["iters"][0].tolist() == d["block_iters"].tolist()
  assert relmax(ns["phi"][0], d["phi"]) < 1e-10 and relmax(ns["rho"][0], d["rho"]) < 1e-10

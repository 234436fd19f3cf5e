"""Builds libpdhg_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python pdhg-optimal-control_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("pdhg_api.cu", "pdhg1d_cta.cu", "pdhg1d_k1.cu", "pdhg_aux.cu", "pdhg_traj.cu", "pdhg_coop.cu")]
HDR = [os.path.join(HERE, "csrc", f) for f in ("pdhg_device.cuh", "pdhg_params.h")] + \
      [os.path.join(os.path.dirname(HERE), "include", "pdhg_b200.h")]
OUT = os.path.join(HERE, "lib", "libpdhg_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


def up_to_date():
  if not os.path.exists(OUT):
    return False
  t = os.path.getmtime(OUT)
  return all(os.path.getmtime(f) <= t for f in SRC + HDR + [os.path.abspath(__file__)])


def build(force=False, verbose=False, defs=(), out=None):
  """`defs` / `out`: experiment builds (extra -D macros into a differently named library, selected with PDHG_B200_LIB)."""
  if out is None and not force and up_to_date():
    return OUT
  out = out or OUT
  os.makedirs(os.path.dirname(out), exist_ok=True)
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defs] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SRC
  r = subprocess.run(cmd, capture_output=True, text=True)
  if verbose or r.returncode != 0:
    sys.stderr.write(r.stdout + r.stderr)
  if r.returncode != 0:
    raise RuntimeError("nvcc failed: " + " ".join(cmd))
  return out


if __name__ == "__main__":
  defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
  outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defs=defs, out=os.path.join(HERE, "lib", outs[0]) if outs else None))

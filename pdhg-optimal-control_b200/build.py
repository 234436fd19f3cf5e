"""Builds libpdhg_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python pdhg-optimal-control_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("pdhg_api.cu", "pdhg1d_cta.cu", "pdhg1d_k1.cu", "pdhg_aux.cu", "pdhg_coop.cu")]
HDR = [os.path.join(HERE, "csrc", f) for f in ("pdhg_device.cuh", "pdhg_params.h")] + \
      [os.path.join(os.path.dirname(HERE), "include", "pdhg_b200.h")]
OUT = os.path.join(HERE, "lib", "libpdhg_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def up_to_date():
  if not os.path.exists(OUT):
    return False
  t = os.path.getmtime(OUT)
  return all(os.path.getmtime(f) <= t for f in SRC + HDR + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
  if not force and up_to_date():
    return OUT
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
  cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
  r = subprocess.run(cmd, capture_output=True, text=True)
  if verbose or r.returncode != 0:
    sys.stderr.write(r.stdout + r.stderr)
  if r.returncode != 0:
    raise RuntimeError("nvcc failed: " + " ".join(cmd))
  return OUT


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

A NumPy-backed stand-in for the *tiny* slice of the `jax` / `einshape` API that the
reference's hot-path sources use, so that the UNMODIFIED reference files under
`/root/reference/jaxsrc` can be executed in the build container (where jax/jaxlib,
einshape, tensorflow, haiku, optax, pytz and matplotlib are not installed and cannot
be).  It exists for exactly one purpose: `oracle/make_golden.py` imports the reference
through it to (a) validate the NumPy restatement in `oracle/pdhg_numpy.py` and
(b) write the golden vectors under `tests/golden/`.

What this is NOT: it is not jaxlib.  The arithmetic is NumPy's (pocketfft FFT, IEEE
double), the control flow and every formula are the reference's own source lines.
XLA-specific rounding (fusion, FMA contraction, its FFT) is therefore not pinned;
everything else about the reference's behaviour is.

API surface covered (reference call sites):
  jax.config.update                      set_fns.py:7, solver.py:11, update_fns_in_pdhg.py:10
  jax.jit (+ static_argnames)            update_fns_in_pdhg.py:135,142,150; utils_diff_op.py:25...
  jax.vmap(in_axes, out_axes)            utils_precond.py:38-40
  jax.lax.scan(f, init, xs, unroll=)     utils_precond.py:19-33
  jax.numpy.* (NumPy namesakes)          everywhere
  jax.scipy.fft.dct / idct               utils_precond.py:52,65,160,174
  einshape.jax_einshape                  utils_precond.py:125,131,167-168; utils_pdhg_solver.py:123
  JAX out-of-range integer index clamp   utils_pdhg_solver.py:90 (`error[2]` on a length-2 array)
"""
import sys
import types

import numpy as np
import scipy.fft as _sfft

REFERENCE_SRC = "/root/reference/jaxsrc"


class _ClampedArray(np.ndarray):
  """ndarray whose scalar integer indexing clamps like JAX's gather (utils_pdhg_solver.py:90)."""

  def __getitem__(self, idx):
    if isinstance(idx, (int, np.integer)) and self.ndim >= 1 and self.shape[0] > 0:
      n = self.shape[0]
      if idx >= n:
        idx = n - 1
      elif idx < -n:
        idx = 0
    out = np.ndarray.__getitem__(self, idx)
    return out


def _array(obj, dtype=None):
  try:
    a = np.array(obj, dtype=dtype)
  except ValueError:
    # Ragged list of per-block error logs (utils_pdhg_solver.py:224 builds jnp.array(errs_all) only to
    # print its max; real JAX raises there too when blocks logged different numbers of rows, i.e. when
    # print_freq is smaller than a block's iteration count).  Tolerated here so that the results computed
    # BEFORE that print can still be captured as golden vectors.
    a = np.concatenate([np.asarray(o, dtype=float).ravel() for o in obj])
  if a.ndim == 1:
    return a.view(_ClampedArray)
  return a


def _jit(fun=None, **_kw):
  if fun is None:
    return lambda f: f
  return fun


def _vmap(fun, in_axes=0, out_axes=0):
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    n = None
    for a, ax in zip(args, axes):
      if ax is not None:
        n = np.shape(a)[ax]
        break
    outs = []
    for i in range(n):
      sl = [a if ax is None else np.take(a, i, axis=ax) for a, ax in zip(args, axes)]
      outs.append(fun(*sl))
    return np.stack(outs, axis=out_axes)
  return mapped


def _scan(f, init, xs, length=None, reverse=False, unroll=1):
  n = len(xs[0]) if isinstance(xs, (tuple, list)) else len(xs)
  carry = init
  ys = []
  for i in range(n):
    x = tuple(a[i] for a in xs) if isinstance(xs, (tuple, list)) else xs[i]
    carry, y = f(carry, x)
    ys.append(y)
  return carry, np.stack(ys, axis=0)


def _einshape(equation, arr, **sizes):
  """The four patterns the reference uses (tile / broadcast)."""
  arr = np.asarray(arr)
  eq = equation.replace(" ", "")
  if eq == "i...->(ki)...":       # utils_pdhg_solver.py:123: k copies of the i-block stacked on axis 0
    return np.concatenate([arr] * sizes["k"], axis=0)
  if eq == "n->mn":               # utils_precond.py:125
    return np.broadcast_to(arr[None, :], (sizes["m"],) + arr.shape).copy()
  if eq == "n->nm":               # utils_precond.py:131
    return np.broadcast_to(arr[:, None], arr.shape + (sizes["m"],)).copy()
  if eq == "n->nmk":              # utils_precond.py:167
    return np.broadcast_to(arr[:, None, None], arr.shape + (sizes["m"], sizes["k"])).copy()
  if eq == "nk->mnk":             # utils_precond.py:168
    return np.broadcast_to(arr[None], (sizes["m"],) + arr.shape).copy()
  raise NotImplementedError(equation)


class _Permissive(types.ModuleType):
  """Module whose unknown attributes are inert placeholders (tensorflow, haiku, optax, ...)."""

  def __getattr__(self, name):
    if name.startswith("__"):
      raise AttributeError(name)
    if name == "Module":
      return object
    sub = _Permissive(self.__name__ + "." + name)
    setattr(self, name, sub)
    return sub

  def __call__(self, *a, **k):
    return None


def install(reference_src=REFERENCE_SRC):
  """Registers the stand-in modules and puts the reference sources on sys.path."""
  if "jax" in sys.modules and getattr(sys.modules["jax"], "__pdhg_shim__", False):
    return
  jnp = types.ModuleType("jax.numpy")
  for name in dir(np):
    if not name.startswith("_"):
      setattr(jnp, name, getattr(np, name))
  jnp.array = _array
  jnp.fft = np.fft
  jnp.linalg = np.linalg
  jnp.complex128 = np.complex128
  jnp.pi = np.pi

  lax = types.ModuleType("jax.lax")
  lax.scan = _scan

  jsp = types.ModuleType("jax.scipy")
  jsp_fft = types.ModuleType("jax.scipy.fft")
  jsp_fft.dct = lambda x, type=2, n=None, axis=-1, norm=None: _sfft.dct(x, type=type, n=n, axis=axis, norm=norm)
  jsp_fft.idct = lambda x, type=2, n=None, axis=-1, norm=None: _sfft.idct(x, type=type, n=n, axis=axis, norm=norm)
  jsp.fft = jsp_fft

  tree = types.ModuleType("jax.tree_util")
  tree.tree_map = lambda f, *t: None
  tree.tree_leaves = lambda t: []

  jax = types.ModuleType("jax")
  jax.__pdhg_shim__ = True
  jax.numpy = jnp
  jax.lax = lax
  jax.scipy = jsp
  jax.tree_util = tree
  jax.jit = _jit
  jax.vmap = _vmap
  jax.config = types.SimpleNamespace(update=lambda *a, **k: None)

  ein = types.ModuleType("einshape")
  ein.jax_einshape = _einshape

  pytz = types.ModuleType("pytz")
  pytz.timezone = lambda name: None

  mods = {"jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.scipy": jsp, "jax.scipy.fft": jsp_fft,
          "jax.tree_util": tree, "einshape": ein, "pytz": pytz}
  for name in ("tensorflow", "haiku", "optax", "matplotlib", "matplotlib.pyplot"):
    mods[name] = _Permissive(name)
  sys.modules.update(mods)
  if reference_src not in sys.path:
    sys.path.insert(0, reference_src)

"""pdhg_b200 — host-side mirror of the reference's solver API over the B200-native C ABI.

Module names follow the reference (`jaxsrc/`): `set_fns`, `update_fns_in_pdhg`, `solver`,
`utils.utils_pdhg_solver`, `utils.utils_precond`, `run_example`.  All arithmetic of the PDHG hot path runs in
`lib/libpdhg_b200.so` (hand-written sm_100a CUDA); importing this package without that library fails.
"""
from . import _lib

_lib.load()   # fail loudly: there is no CPU fallback

from .set_fns import set_up_example_fns, set_up_J, Functions          # noqa: E402,F401
from .solver import save, load_solution, load_middle_solution         # noqa: E402,F401
from .utils.utils_pdhg_solver import PDHG_multi_step, PDHG_solver_oneiter   # noqa: E402,F401
from .run_example import solve_HJ                                     # noqa: E402,F401

"""Diagnostic: host enqueue time against device time of the pieces of a slab iteration (torchrun, one rank per GPU)."""
import contextlib, io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pdhg-optimal-control_b200"))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from pdhg_b200 import run_example as rx, set_fns as sf, slab
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x_arr = rx.make_x_arr(2, nx, nx, 2.0, 2.0)
with contextlib.redirect_stdout(io.StringIO()):
  fns = sf.set_up_example_fns(1, 2, 0)
g = sf.set_up_J(1, 2, (2.0, 2.0))(x_arr)[0]
R, grp, kind = slab.make_dist_rank(rank, world, dist, fns, nx, nx, 1.0 / 256, (2.0 / nx, 2.0 / nx), 70.0, x_arr, device=local)
slab.init_block(grp, g, 70.0)
slab.solve_block_slab(grp, 0.1, 5e-4, 3)
sigma, epsl = 5e-4 * 1.5, 0.1
pieces = {
  "A": lambda: R.ext(R.hL, 0, 0.0, epsl, rho_in=R.rho[0], alp_in=R.alp[0], zt=R.zt),
  "B": lambda: R.ext(R.hB, 1, 0.0, epsl, zt=R.ztB, nyh_override=R.kyn, ky_off=R.ky0, nyh_tab=R.nyh),
  "C": lambda: R.ext(R.hL, 2, 1e-4, epsl, zt=R.zt, phi_in=R.phi[0], phi_out=R.phi[1], phib=R.phib),
  "D1": lambda: R.ext(R.hL, 3, sigma, epsl, pass_mask=1, phib=R.phib, rho_in=R.rho[0], alp_in=R.alp[0], rho_out=R.rho[1], alp_out=R.alp[1], sums=R.sums[0]),
  "D2": lambda: R.ext(R.hL, 3, sigma, epsl, pass_mask=2, phib=R.phib, rho_in=R.rho[0], alp_in=R.alp[0], rho_out=R.rho[1], alp_out=R.alp[1], sums=R.sums[0]),
  "D%d" % R.fuse: lambda: R.ext(R.hL, 3, sigma, epsl, pass_mask=R.fuse, phib=R.phib, rho_in=R.rho[0], alp_in=R.alp[0], rho_out=R.rho[1], alp_out=R.alp[1], sums=R.sums[0]),
  "E": lambda: R.ext(R.hL, 4, 0.0, epsl, rho_in=R.rho[0], alp_in=R.alp[0], rho_out=R.rho[1], alp_out=R.alp[1], sums=R.sums[1]),
  "empty_launch": lambda: R.ext(R.hL, 5, 0.0, epsl),
  "halo_dual": lambda: grp.halo(lambda r: [r.dual[0][0:3]]),
  "halo_phib": lambda: grp.halo(lambda r: [r.phib]),
  "transpose_fwd": grp.transpose_fwd,
  "transpose_bwd": grp.transpose_bwd,
  "sync": grp.sync,
  "sums4_incl_host_read": lambda: grp.allreduce_sums(4),
}
out = {"P": world, "exchange": kind, "nx": nx, "fuse": R.fuse}
n = 50
for name, fn in pieces.items():
  for _ in range(3): fn()
  dist.barrier(); torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(n): fn()
  t1 = time.perf_counter()
  torch.cuda.synchronize()
  t2 = time.perf_counter()
  out[name] = {"host_us": round((t1 - t0) / n * 1e6, 1), "total_us": round((t2 - t0) / n * 1e6, 1)}
if rank == 0:
  print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()

"""Diagnostic: seconds per section of the slab iteration (PDHG_SLAB_PROF=1), for a given fuse depth.  torchrun, one rank per GPU."""
import contextlib, io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pdhg-optimal-control_b200"))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from pdhg_b200 import run_example as rx, set_fns as sf, slab
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
x_arr = rx.make_x_arr(2, nx, nx, 2.0, 2.0)
with contextlib.redirect_stdout(io.StringIO()):
  fns = sf.set_up_example_fns(1, 2, 0)
g = sf.set_up_J(1, 2, (2.0, 2.0))(x_arr)[0]
quick = os.environ.get("PDHG_SLAB_PROF_QUICK") is not None
for fuse in ((5,) if quick else (2, 5)):
  os.environ["PDHG_SLAB_FUSE"] = str(fuse)
  R, grp, kind = slab.make_dist_rank(rank, world, dist, fns, nx, nx, 1.0 / 256, (2.0 / nx, 2.0 / nx), 70.0, x_arr, device=local)
  for prof in ((False,) if quick else (False, True)):
    if prof: os.environ["PDHG_SLAB_PROF"] = "1"
    else: os.environ.pop("PDHG_SLAB_PROF", None)
    slab.PROFILE.clear()
    slab.init_block(grp, g, 70.0)
    slab.solve_block_slab(grp, 0.1, 5e-4, 3)
    slab.PROFILE.clear()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    res = slab.solve_block_slab(grp, 0.1, 5e-4, iters)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
      print(json.dumps({"P": world, "exchange": kind, "fused_transposes": [grp.fused_fwd, grp.fused_bwd], "fuse_asked": fuse, "fuse": R.fuse, "prof": prof, "ms_per_iter": dt / iters * 1e3, "n_inner": res[4],
                        "sections_ms_per_iter": {k: round(v / iters * 1e3, 3) for k, v in slab.PROFILE.items()}}), flush=True)
dist.destroy_process_group()

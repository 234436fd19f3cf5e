"""x-slab decomposition of ONE 2-D grid over the ranks of a torchrun launch (BASELINE configs[4] building block):
parity of the NCCL path against the single-GPU kernel + iterations/s of both.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/slab_bench.py [nx] [iters]

Rank r owns x-rows [r nx/P, (r+1) nx/P) of the block (time_step_per_PDHG = 2).  Every outer iteration does two halo exchanges
(batched isend/irecv), two all-to-all transposes of the half spectrum and one small all-reduce per dual sweep
(pdhg_b200/slab.py).  Prints ONE JSON line on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pdhg-optimal-control_b200"))

import numpy as np
import torch
import torch.distributed as dist


def main():
  nx = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
  iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
  rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
  os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
  os.environ.setdefault("MASTER_PORT", "29511")
  torch.cuda.set_device(local)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
  import contextlib, io
  from pdhg_b200 import run_example as rx, set_fns as sf, slab
  ny, T, epsl, stepsz = nx, 1.0 / 256, 0.0, 0.1
  n_ctrl, bc, _ = rx.problem_setup(1, 2)
  x_arr = rx.make_x_arr(2, nx, ny, 2.0, 2.0)
  with contextlib.redirect_stdout(io.StringIO()):
    fns = sf.set_up_example_fns(1, 2, 0)
  g = sf.set_up_J(1, 2, (2.0, 2.0))(x_arr)[0]
  R, grp, kind = slab.make_dist_rank(rank, world, dist, fns, nx, ny, T, (2.0 / nx, 2.0 / ny), 70.0, x_arr, device=local)
  out = {}
  for label, n in (("warm", 3), ("timed", iters)):
    slab.init_block(grp, g, 70.0)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = slab.solve_block_slab(grp, epsl, stepsz, n)
    torch.cuda.synchronize(); dist.barrier()
    out[label] = (time.perf_counter() - t0, res)
  t_slab = torch.tensor([out["timed"][0]], dtype=torch.float64, device="cuda")
  dist.all_reduce(t_slab, op=dist.ReduceOp.MAX)
  phi, rho, alp = slab.gather_block(grp)          # this rank's slab
  # single-GPU solve of the same block on rank 0 (same iteration cap), for parity and as the 1-GPU time
  line = None
  if rank == 0:
    info = {}
    with contextlib.redirect_stdout(io.StringIO()):
      rx.solve_HJ(2, n_ctrl, 1, epsl, fns, nx, ny, 2, 2.0, 2.0, T, x_arr, 70.0, 2, stepsz, 3, 10 ** 9, 1e-6, bc)
      torch.cuda.synchronize()
      t0 = time.perf_counter()
      res1, _ = rx.solve_HJ(2, n_ctrl, 1, epsl, fns, nx, ny, 2, 2.0, 2.0, T, x_arr, 70.0, 2, stepsz, iters, 10 ** 9, 1e-6, bc, info=info)
      torch.cuda.synchronize()
      t1 = time.perf_counter() - t0
    _, phi1, rho1, alp1 = res1[0]
    nxl = nx // world
    rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    par = {"phi": rel(phi, np.asarray(phi1)[:, :nxl]), "rho": rel(rho, np.asarray(rho1)[:, :nxl]),
           "alp": rel(alp, np.stack(alp1)[:, :, :nxl])}
    it_s, it_1 = out["timed"][1][0], info["block_iters"][0]
    line = {"workload": "egno=1 ndim=2 epsl=0 nx=ny=%d, one time block (time_step_per_PDHG=2), x-slab decomposed over %d GPU(s)" % (nx, world),
            "n_gpus": world, "iters": it_s, "iters_single_gpu": it_1, "n_inner": out["timed"][1][4],
            "slab_seconds": float(t_slab[0]), "slab_iters_per_s": it_s / float(t_slab[0]),
            "single_gpu_seconds_incl_h2d_d2h": t1, "single_gpu_iters_per_s": it_1 / t1,
            "parity_rel_linf_rank0_slab_vs_single_gpu": par,
            "exchange": kind,
            "exchanges_per_iteration": "2 halo exchanges, 2 transposes of the half spectrum, 1 sum all-reduce per dual pass (symm: peer stores + "
                                       "device barrier over NVLink; nccl: batch_isend_irecv / all_to_all_single / all_reduce)",
            "note": "host-driven phases (one launch per phase, exit tests on the host after each all-reduce)"}
    print(json.dumps(line), flush=True)
  dist.barrier()
  dist.destroy_process_group()


if __name__ == "__main__":
  main()

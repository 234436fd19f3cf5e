"""Multi-GPU use of the PDHG path: batches of independent problem instances sharded across ranks.

One process per GPU (torch.distributed, NCCL on GPUs / gloo in CPU tests).  Instances are independent, so the
solve itself needs NO collective: each rank marches its contiguous range of instances on its own device; only
the small per-instance logs (iteration counts, status, final step size) are gathered at the end.
"""
import numpy as np


def shard_range(n_items, rank, world_size):
  """Contiguous, balanced [begin, end) of `n_items` for `rank` (first n_items % world ranks get one extra)."""
  base, extra = divmod(int(n_items), int(world_size))
  begin = rank * base + min(rank, extra)
  return begin, begin + base + (1 if rank < extra else 0)


def gather_instance_logs(local, n_items, dist=None):
  """All-gathers per-instance 1-D log arrays (dict name -> array over the local shard) into full-length arrays.
  Uses fixed-size padded all_gather (works for NCCL with CUDA tensors and gloo with CPU tensors)."""
  if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
    return {k: np.asarray(v) for k, v in local.items()}
  import torch
  world, rank = dist.get_world_size(), dist.get_rank()
  cap = max(shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world))
  dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
  out = {}
  for name, arr in local.items():
    arr = np.asarray(arr, dtype=np.float64).ravel()
    buf = torch.zeros(cap, dtype=torch.float64, device=dev)
    buf[:arr.size] = torch.from_numpy(arr).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    full = np.empty(n_items, dtype=np.float64)
    for r in range(world):
      b, e = shard_range(n_items, r, world)
      full[b:e] = parts[r][: e - b].cpu().numpy()
    out[name] = full
  return out


def solve_batch_sharded(solve_fn, g, epsl, stepsz, rank=0, world_size=1, dist=None):
  """Runs `solve_fn(g_shard, epsl_shard, stepsz_shard) -> (phi, rho, alp, logs)` on this rank's instances and
  gathers the per-instance logs.  Returns (begin, end, phi, rho, alp, gathered_logs)."""
  B = len(g)
  b, e = shard_range(B, rank, world_size)
  epsl = np.broadcast_to(np.asarray(epsl, dtype=np.float64), (B,))
  stepsz = np.broadcast_to(np.asarray(stepsz, dtype=np.float64), (B,))
  phi, rho, alp, logs = solve_fn(g[b:e], epsl[b:e], stepsz[b:e])
  local = {"total_iters": logs.iters.sum(axis=1), "max_iters": logs.iters.max(axis=1), "status": logs.status,
           "stepsz_final": logs.stepsz_final, "blocks_done": logs.blocks_done}
  return b, e, phi, rho, alp, gather_instance_logs(local, B, dist)

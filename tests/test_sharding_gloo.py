"""CPU, world_size 2 over gloo: the instance-sharding host logic of the batched (cfg4-style) multi-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pdhg_b200.sharding import gather_instance_logs, shard_range, solve_batch_sharded


def test_shard_range_partitions():
  for B in (1, 2, 7, 4096, 4099):
    for W in (1, 2, 3, 8):
      ranges = [shard_range(B, r, W) for r in range(W)]
      assert ranges[0][0] == 0 and ranges[-1][1] == B
      assert all(ranges[i][1] == ranges[i + 1][0] for i in range(W - 1))
      sizes = [e - b for b, e in ranges]
      assert max(sizes) - min(sizes) <= 1


class _FakeLogs:
  def __init__(self, idx, nblocks=3):
    self.iters = np.stack([np.arange(nblocks) + 10 * i for i in idx]).astype(np.int64)
    self.status = np.array([i % 2 for i in idx], np.int32)
    self.stepsz_final = np.array([0.1 / (1 + i) for i in idx])
    self.blocks_done = np.full(len(idx), nblocks, np.int32)


def _worker(rank, world, port, B, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  g = np.arange(B, dtype=np.float64)[:, None] * np.ones((1, 4))

  def fake_solve(gs, es, ss):   # stands in for run_example.solve_HJ_batch on this rank's device
    idx = gs[:, 0].astype(int)
    return gs * 2, gs * 3, gs * 4, _FakeLogs(idx)

  b, e, phi, rho, alp, logs = solve_batch_sharded(fake_solve, g, 0.0, 0.1, rank, world, dist)
  ok = (b, e) == shard_range(B, rank, world) and np.array_equal(phi, g[b:e] * 2)
  ok &= np.array_equal(logs["total_iters"], np.array([3 + 30 * i for i in range(B)], dtype=float))
  ok &= np.array_equal(logs["status"], np.arange(B) % 2)
  ok &= np.allclose(logs["stepsz_final"], 0.1 / (1 + np.arange(B)))
  q.put((rank, bool(ok)))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_solve_gathers_logs_world2():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
  for p in procs: p.start()
  res = sorted(q.get(timeout=100) for _ in range(2))
  for p in procs: p.join(timeout=30)
  assert res == [(0, True), (1, True)]


def test_gather_without_dist_is_identity():
  out = gather_instance_logs({"a": np.arange(5)}, 5, None)
  assert np.array_equal(out["a"], np.arange(5))

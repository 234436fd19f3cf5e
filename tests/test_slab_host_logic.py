"""CPU: the exchange steps of the x-slab decomposition (pdhg_b200/slab.py) on CPU tensors — no kernels involved.

* `LocalGroup` (P emulated ranks): the two transposes against a NumPy statement of what they must do, and halo rows.
* `DistGroup` over gloo, world_size 2 and 3: halo exchange (incl. the P = 2 case where left == right), all-to-all transposes
  and the sum all-reduce give every rank exactly what `LocalGroup` gives it."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pdhg_b200 import slab


def _fake_ranks(P, nx, ny, K=1, seed=0):
  """Rank objects with the fields the group classes touch, filled with recognisable data (CPU tensors)."""
  rng = np.random.default_rng(seed)
  nxl, nyh = nx // P, ny // 2 + 1
  kyl = (nyh + P - 1) // P
  zt_glob = rng.standard_normal((K, nyh, nx)) + 1j * rng.standard_normal((K, nyh, nx))
  rho_glob = rng.standard_normal((K, nx, ny))
  ranks = []
  for r in range(P):
    R = SimpleNamespace(rank=r, P=P, K=K, nx=nx, ny=ny, nxl=nxl, nxp=nxl + 2, nyh=nyh, kyl=kyl, dev=torch.device("cpu"))
    R.ky0 = min(r * kyl, nyh)
    R.kyn = max(0, min(kyl, nyh - R.ky0))
    R.zt = torch.zeros((K, nyh, nxl + 2), dtype=torch.complex128)
    R.zt[:, :, 1:nxl + 1] = torch.from_numpy(zt_glob[:, :, r * nxl:(r + 1) * nxl])
    R.ztB = torch.zeros((K, max(kyl, 1), nx), dtype=torch.complex128)
    R.rho = torch.zeros((K, nxl + 2, ny), dtype=torch.float64)
    R.rho[:, 1:nxl + 1] = torch.from_numpy(rho_glob[:, r * nxl:(r + 1) * nxl])
    R.sums = torch.arange(slab.NQ, dtype=torch.float64) * (r + 1)
    ranks.append(R)
  return ranks, zt_glob, rho_glob


@pytest.mark.parametrize("P,nx,ny,K", [(1, 8, 6, 1), (2, 8, 6, 1), (3, 12, 10, 2), (4, 16, 8, 1), (4, 8, 32, 1)])
def test_local_group_transposes_and_halo(P, nx, ny, K):
  ranks, zt_glob, rho_glob = _fake_ranks(P, nx, ny, K)
  grp = slab.LocalGroup(ranks)
  grp.transpose_fwd()
  for R in ranks:        # rank r now owns ky rows [ky0, ky0 + kyn) of the full-x spectrum
    assert np.array_equal(R.ztB[:, :R.kyn].numpy(), zt_glob[:, R.ky0:R.ky0 + R.kyn])
  for R in ranks:        # something rank-specific happens in phase B; the way back must return it to the x-slabs
    R.ztB *= (R.rank + 2)
    R.zt.zero_()
  grp.transpose_bwd()
  owner = np.minimum(np.arange(ranks[0].nyh) // ranks[0].kyl, P - 1)
  for R in ranks:
    want = zt_glob[:, :, R.rank * R.nxl:(R.rank + 1) * R.nxl] * (owner + 2)[None, :, None]
    assert np.array_equal(R.zt[:, :, 1:R.nxl + 1].numpy(), want)
  grp.halo(lambda R: [R.rho])
  for R in ranks:
    lo = R.rank * R.nxl
    assert np.array_equal(R.rho[:, 0].numpy(), rho_glob[:, (lo - 1) % nx])
    assert np.array_equal(R.rho[:, R.nxl + 1].numpy(), rho_glob[:, (lo + R.nxl) % nx])
  assert np.array_equal(grp.allreduce_sums(), np.arange(slab.NQ) * (P * (P + 1) / 2))


def _worker(rank, world, port, nx, ny, K, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  ok = True
  try:
    ref, _, _ = _fake_ranks(world, nx, ny, K)
    lg = slab.LocalGroup(ref)
    mine, _, _ = _fake_ranks(world, nx, ny, K)
    R = mine[rank]
    dg = slab.DistGroup(R, dist)
    lg.halo(lambda r: [r.rho]); dg.halo(lambda r: [r.rho])
    ok &= bool(torch.equal(R.rho, ref[rank].rho))
    lg.transpose_fwd(); dg.transpose_fwd()
    ok &= bool(torch.equal(R.ztB, ref[rank].ztB))
    for r in ref:
      r.ztB *= (r.rank + 2)
    R.ztB *= (rank + 2)
    lg.transpose_bwd(); dg.transpose_bwd()
    ok &= bool(torch.equal(R.zt, ref[rank].zt))
    ok &= bool(np.array_equal(dg.allreduce_sums(), lg.allreduce_sums()))
  except Exception as ex:      # report instead of hanging the peer
    ok = False
    q.put((rank, repr(ex)))
  q.put((rank, bool(ok)))
  dist.barrier()
  dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,nx,ny,K", [(2, 8, 6, 1), (3, 12, 10, 2)])
def test_dist_group_matches_local_group_gloo(world, nx, ny, K):
  s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, K, q)) for r in range(world)]
  for p in procs: p.start()
  res = [q.get(timeout=150) for _ in range(world)]
  for p in procs: p.join(timeout=30)
  assert all(ok is True for _, ok in res), res

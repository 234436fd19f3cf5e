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


def test_exit_test_of_a_fused_pass_matches_the_scalar_rule():
  """slab._first_hit / _last_sweep_sums against the sweep-by-sweep statement of update_fns_in_pdhg.py:170-177 on random totals,
  incl. NaN, zero denominators and several sweeps below eps (the FIRST one counts)."""
  rng = np.random.default_rng(3)

  def scalar(v, ns, eps):
    for sw in range(ns):
      b0 = 0 if sw == 0 else 20 + 16 * (sw - 1)
      with np.errstate(all="ignore"):
        err = v[b0] / v[b0 + 1] + sum(v[b0 + 2 + 2 * q] / v[b0 + 3 + 2 * q] for q in range(4))
      if err < eps:
        return sw
    return -1

  hits = set()
  for trial in range(400):
    v = rng.random(slab.NQ) + 0.5
    for sw in range(5):                       # numerators small in some sweeps so that the test fires there
      b0 = 0 if sw == 0 else 20 + 16 * (sw - 1)
      if rng.random() < 0.3:
        v[b0:b0 + 10:2] *= 1e-9
    if trial % 7 == 0:
      v[rng.integers(0, slab.NQ)] = np.nan
    if trial % 11 == 0:
      v[2 * rng.integers(0, 5) + 1] = 0.0
    for ns in range(1, 6):
      got = slab._first_hit(v, ns, 1e-6)
      assert got == scalar(v, ns, 1e-6), (trial, ns)
      hits.add(got)
      w = slab._last_sweep_sums(v, ns)
      want = v[:16] if ns == 1 else v[20 + 16 * (ns - 2):36 + 16 * (ns - 2)]
      assert np.array_equal(w[:16], want, equal_nan=True) and np.array_equal(w[16:], v[16:], equal_nan=True)
  assert hits >= {-1, 0, 1, 2, 3, 4}


class _ScriptedRank(SimpleNamespace):
  """Stands in for SlabRank on the CPU: `ext` records the launches and writes scripted error sums, so that the control flow of
  solve_block_slab (speculative pass plan, replay of the exit tests, pass-by-pass redo, buffer rotation) runs without a kernel.
  The 'dual sweep' is x -> x / 2 on a scalar state per dual buffer, the relative change of sweep s of an outer iteration is
  script[iteration][s] (below eps = exit)."""

  def interior(self, a):
    return a[..., 1:self.nxl + 1, :]

  def ext(self, h, phase, step, epsl, **kw):
    self.calls.append((phase, kw.get("pass_mask")))
    if phase == 2:
      self.outer += 1
      return
    if phase == 3:
      ns, row = kw["pass_mask"], kw["sums"]
      src = next(b for b in range(3) if kw["rho_in"] is self.rho[b])
      dst = next(b for b in range(3) if kw["rho_out"] is self.rho[b])
      j0 = self.sweeps_of[src] if src != self.cd else 0
      errs = self.script[min(self.outer, len(self.script) - 1)]
      row.zero_()
      for s in range(ns):
        b0 = 0 if s == 0 else 20 + 16 * (s - 1)
        e = errs[min(j0 + s, len(errs) - 1)]
        row[b0] = e; row[b0 + 1] = 1.0
        for q in range(4):
          row[b0 + 2 + 2 * q] = 0.0; row[b0 + 3 + 2 * q] = 1.0
      row[16] = 1.0; row[17] = 1.0                      # err1 = 1 / sqrt(S + 1): never converges on its own
      self.sweeps_of[dst] = j0 + ns
      self.state[dst] = self.state[src] / 2.0 ** ns
    if phase == 4:
      kw["sums"].zero_()
      kw["sums"][10] = 1.0


def _scripted_solve(script, n_iter, spec, fuse):
  os.environ["PDHG_SLAB_SPEC"] = spec
  try:
    nxl, ny = 4, 4
    z = lambda *sh: torch.zeros(sh, dtype=torch.float64)
    R = _ScriptedRank(rank=0, P=1, K=1, nx=nxl, ny=ny, nxl=nxl, nxp=nxl + 2, nyh=ny // 2 + 1, kyl=ny // 2 + 1, ky0=0, kyn=ny // 2 + 1,
                      dev=torch.device("cpu"), fuse=fuse, hL=None, hB=None, cp=0, cd=0, calls=[], outer=-1, script=script,
                      sweeps_of=[0, 0, 0], state=[1.0, 0.0, 0.0])
    R.phi = [z(2, nxl + 2, ny) + 1.0, z(2, nxl + 2, ny)]
    R.phib = z(2, nxl + 2, ny)
    R.dual = [z(5, 1, nxl + 2, ny) + 1.0 for _ in range(3)]
    R.rho = [d[0] for d in R.dual]
    R.alp = [d[1:5] for d in R.dual]
    R.zt = torch.zeros((1, R.nyh, nxl + 2), dtype=torch.complex128)
    R.ztB = torch.zeros((1, R.kyl, nxl), dtype=torch.complex128)
    R.sums = z(slab.NSLOT, slab.NQ)
    grp = slab.LocalGroup([R], fused="0")
    res = slab.solve_block_slab(grp, 0.0, 0.1, n_iter)
    return res, dict(slab.STATS), R
  finally:
    os.environ.pop("PDHG_SLAB_SPEC", None)


@pytest.mark.parametrize("fuse", [1, 2, 4, 5])
def test_speculative_pass_plan_control_flow_on_scripted_sums(fuse):
  """Sweep counts that fall, stay and rise again from one outer iteration to the next: the speculative plan (all passes enqueued,
  one exchange, replay on the host) must do exactly the sweeps of the pass-by-pass loop, end on the same dual buffer with the same
  state, and fall back to the pass-by-pass loop precisely when the prediction was wrong."""
  big, hit = 1.0, 1e-9
  script = [[big] * 10,                        # 10 sweeps, no exit
            [big] * 6 + [hit],                 # exit at sweep 7  (predicted 10: redo)
            [big] * 6 + [hit],                 # exit at sweep 7  (predicted 7: the plan holds)
            [big] * 2 + [hit],                 # exit at sweep 3  (predicted 7: redo)
            [big] * 8 + [hit],                 # exit at sweep 9  (predicted 3: the plan ends early: redo)
            [hit],                             # exit at sweep 1  (predicted 9: redo)
            [hit]]                             # exit at sweep 1  (predicted 1: the plan holds)
  want_inner = 10 + 7 + 7 + 3 + 9 + 1 + 1
  out = {}
  for spec in ("0", "1"):
    res, stats, R = _scripted_solve(script, len(script), spec, fuse)
    out[spec] = (res, stats, R)
    assert res[0] == len(script) and res[4] == want_inner, (spec, res)
    assert R.state[R.cd] == 2.0 ** -want_inner           # the accepted dual buffer went through exactly the sweeps that count
  assert out["0"][1]["respeculated"] == 0
  assert out["1"][1]["respeculated"] == 4        # iterations 1, 3, 4 and 5: the sweep count differs from the previous iteration's
  assert out["0"][0] == out["1"][0] and out["0"][2].cd == out["1"][2].cd

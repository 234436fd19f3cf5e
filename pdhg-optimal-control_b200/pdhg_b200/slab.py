"""x-slab decomposition of ONE large 2-D grid over P GPUs (BASELINE configs[4]).

Every rank owns nx/P consecutive x-rows (all y, all t-rows of the block) plus one ghost row on each side.  One outer PDHG
iteration (jaxsrc/utils/utils_pdhg_solver.py:51-88) becomes
    halo(rho, alp1_x, alp2_x) -> A (residual + y-FFT) -> all-to-all transpose (ky-slabs) -> B (x-FFT, per-mode solve,
    inverse x-FFT) -> all-to-all back -> C (inverse y-FFT, phi update) -> halo(phi_bar) -> D (dual sweeps), each sweep
    followed by ONE small all-reduce of the 20 error sums; the exit tests of the reference run on the host on those sums.
The phases are the cooperative kernel's own (`pdhg_ext_phase`, caller-owned buffers); the exchange steps are
torch.distributed collectives (NCCL on GPUs): `all_to_all_single` for the FFT transposes, `all_reduce` for the sums,
batched point-to-point sends for the halo rows.  `LocalGroup` runs the same algorithm with P emulated ranks inside one
process on one GPU (tensor copies instead of collectives) — that is how the decomposition is parity-tested without a
multi-GPU box.  Restrictions of this first version: 2-D, periodic, time_step_per_PDHG = 2 (K = 1), nx divisible by P.
"""
import os
import time

import numpy as np

from . import _dev, _lib
from .set_fns import coef_tables

NQ = 84      # totals a phase launch hands out: 0..15 dual sweep, 16..18 primal, 20 + 16 (s - 2) ..: sweep s >= 2 of a fused pass
NSLOT = 12   # rows of `sums` per rank: one per launch of an iteration's dual loop (<= 10 sweep passes + the outer-error pass)
FUSE = 5     # dual sweeps per pass while the inner loop is long, capped by what the handle can fuse (pdhg_max_fuse)


class SlabRank:
  """Buffers and handles of one rank."""

  def __init__(self, rank, P, fns_dict, nx, ny, dt, dspatial, c_on_rho, x_arr, C=1.0, eps=1e-6, device=0, arena_alloc=None):
    """`arena_alloc(n_doubles) -> fp64 tensor`: where the exchanged buffers live.  Default: ordinary device memory; SymmGroup passes
    torch's symmetric-memory allocator so that every rank's buffers are peer-mapped over NVLink with the SAME layout."""
    t = _dev.require_cuda()
    assert fns_dict.ndim == 2 and nx % P == 0 and ny % 2 == 0
    self.rank, self.P, self.nx, self.ny, self.K = rank, P, nx, ny, 1
    self.nxl = nx // P
    self.nxp = self.nxl + 2
    self.nyh = ny // 2 + 1
    self.kyl = (self.nyh + P - 1) // P
    self.ky0 = min(rank * self.kyl, self.nyh)
    self.kyn = max(0, min(self.kyl, self.nyh - self.ky0))      # ky rows this rank owns after the transpose
    self.dev = t.device("cuda", device)
    self.eps = eps
    cx, cy = coef_tables(fns_dict.egno, 2, np.asarray(x_arr))
    i0 = rank * self.nxl
    idx = (np.arange(-1, self.nxl + 1) + i0) % nx              # global x index of every local (ghost-padded) row
    self.i0 = i0
    dx, dy = float(dspatial[0]), float(dspatial[1])
    # local handle: the padded slab treated as a small periodic problem (its wrap only touches the ghost rows)
    self.hL = _lib.Solver(2, fns_dict.egno, self.nxp, ny, 1, fns_dict.n_ctrl, (0, 0), float(dt), dx, dy, float(c_on_rho), cx[idx], cy,
                          float(C), 1.0, 1.0, float(eps), 10, 1, 1, 4, device, 2)
    # global handle: owns the per-mode table of the full grid; phase B runs on this rank's ky-slab of it.  path 3 = tables only:
    # no state and no workspace of the full grid on this rank (a rank's memory must shrink with P)
    self.hB = _lib.Solver(2, fns_dict.egno, nx, ny, 1, fns_dict.n_ctrl, (0, 0), float(dt), dx, dy, float(c_on_rho), cx, cy,
                          float(C), 1.0, 1.0, float(eps), 10, 1, 1, 4, device, 3)
    f64 = t.float64
    # every buffer another rank reads or writes is carved from ONE arena (identical layout on all ranks)
    nxp, nyh, kyl = self.nxp, self.nyh, max(self.kyl, 1)
    sizes = [("phi0", 2 * nxp * ny), ("phi1", 2 * nxp * ny), ("phib", 2 * nxp * ny), ("dual0", 5 * nxp * ny), ("dual1", 5 * nxp * ny),
             ("dual2", 5 * nxp * ny), ("zt", 2 * nyh * nxp), ("ztB", 2 * kyl * nx), ("sums_all", 2 * P * NSLOT * NQ)]
    total = sum(n for _, n in sizes)
    self.arena = arena_alloc(total) if arena_alloc else t.zeros(total, dtype=f64, device=self.dev)
    self.arena.zero_()
    off, o = {}, 0
    for name, n in sizes:
      off[name] = o
      o += n
    self.off = off
    carve = lambda name, *sh: self.arena[off[name]:off[name] + int(np.prod(sh))].view(*sh)
    self.phi = [carve("phi0", 2, nxp, ny), carve("phi1", 2, nxp, ny)]
    self.phib = carve("phib", 2, nxp, ny)
    # the outer iterate + two work buffers of the inner dual loop; rho and the four control arrays of a buffer share ONE tensor
    # [5][1][nxp][ny], so the ghost rows of (rho, alp1_x, alp2_x) are a single strided slice: one pack, one unpack kernel
    self.dual = [carve("dual%d" % b, 5, 1, nxp, ny) for b in range(3)]
    self.rho = [d[0] for d in self.dual]
    self.alp = [d[1:5] for d in self.dual]
    self.ztpad = None                                        # send buffer of the forward transpose (padding rows stay zero)
    self.zt = t.view_as_complex(carve("zt", 1, nyh, nxp, 2))
    self.ztB = t.view_as_complex(carve("ztB", 1, kyl, nx, 2))
    self.sums_all = carve("sums_all", 2, P, NSLOT * NQ)       # SymmGroup: every rank's sums, two alternating sets
    z = lambda *sh: t.zeros(sh, dtype=f64, device=self.dev)
    self.sums = z(NSLOT, NQ)                                  # one row of grid totals per launch of an iteration's dual loop
    self.cp, self.cd = 0, 0
    self.fuse = max(1, min(int(os.environ.get("PDHG_SLAB_FUSE", FUSE)), self.hL.max_fuse))
    self._stream = _dev.stream_ptr(self.dev.index)

  def interior(self, a):
    return a[..., 1:self.nxl + 1, :]

  def ext(self, h, phase, step, epsl, **kw):
    scalars = ("nyh_override", "ky_off", "nyh_tab", "pass_mask")
    ptr = {k: (v.data_ptr() if v is not None else None) for k, v in kw.items() if k not in scalars}
    extra = {k: v for k, v in kw.items() if k in scalars}
    h.ext_phase(phase, step, epsl, 1, self.nxl + 1, stream=self._stream, **extra, **ptr)


def _fused_default():
  """Which transposes are fused into the producing / consuming kernel (pdhg_ext_set_exchange; env PDHG_SLAB_FUSED_XCH):
    "B"    (default) phase B GATHERS the rows of its ky-slab straight from the owners' x-slabs and SCATTERS its result straight back:
           both transposes live inside one kernel, no ky-slab buffer, no copy kernels - only the two barriers remain;
    "bwd"  phase B scatters only (the forward transpose stays a set of strided peer copies);
    "both" phase A scatters its half spectrum into the owners' ky-slabs and phase B scatters back (A's transposed store reaches a remote
           slab in 32-96-byte pieces, which slows A by more than the copy it saves: 2048^2, P = 2: A 89 -> 121 us for 45 -> 26 us);
    "0"    strided peer copies both ways."""
  return os.environ.get("PDHG_SLAB_FUSED_XCH", "B")


def _set_exchange(ranks, fwd_ptr, bwd_ptr, mode):
  """`bwd_ptr(R, d)` = rank d's zt as seen from R's device, `fwd_ptr(R, d)` = rank d's ky-slab buffer ztB.  Returns (fused_fwd,
  fused_bwd); (False, False) if the handles' transforms cannot do it (warp-private 256-point variants)."""
  mode = {True: "B", False: "0", None: "0", "1": "B"}.get(mode, mode)
  if mode not in ("B", "bwd", "both") or not all(hasattr(R, "hL") and R.hL.exchange_ok and R.hB.exchange_ok for R in ranks):
    return False, False
  for R in ranks:
    if mode == "both":
      R.hL.set_exchange(R.P, R.rank, R.nxl, R.kyl, R.nyh, [fwd_ptr(R, d) for d in range(R.P)])
    R.hB.set_exchange(R.P, R.rank, R.nxl, R.kyl, R.nyh, [bwd_ptr(R, d) for d in range(R.P)], gather=(mode == "B"))
  return mode in ("both", "B"), True


class LocalGroup:
  """P emulated ranks in one process / on one GPU: collectives are tensor copies (used for parity tests)."""

  def __init__(self, ranks, fused=None):
    self.ranks = ranks
    self.P = len(ranks)
    self.fused_fwd, self.fused_bwd = _set_exchange(ranks, lambda R, d: ranks[d].ztB.data_ptr(), lambda R, d: ranks[d].zt.data_ptr(),
                                                   _fused_default() if fused is None else fused)

  def halo(self, get, fenced=False):
    P = self.P
    for r, R in enumerate(self.ranks):
      for a, left, right in zip(get(R), get(self.ranks[(r - 1) % P]), get(self.ranks[(r + 1) % P])):
        a[..., 0, :] = left[..., R.nxl, :]
        a[..., R.nxl + 1, :] = right[..., 1, :]

  def transpose_fwd(self):
    if self.fused_fwd:        # phase A stored its spectrum straight into the ranks' ky-slabs
      return
    t = _dev.torch()
    P, R0 = self.P, self.ranks[0]
    K, kyl, nxl, nyh = R0.K, R0.kyl, R0.nxl, R0.nyh
    send = []
    for R in self.ranks:
      z = t.zeros((K, P * kyl, nxl), dtype=R.zt.dtype, device=R.dev)
      z[:, :nyh] = R.zt[:, :, 1:nxl + 1]
      send.append(z.view(K, P, kyl, nxl).permute(1, 0, 2, 3).contiguous())          # [dest][K][kyl][nxl]
    for d, R in enumerate(self.ranks):
      recv = t.stack([send[s][d] for s in range(P)], dim=0)                           # [src][K][kyl][nxl]
      R.ztB.copy_(recv.permute(1, 2, 0, 3).reshape(K, kyl, P * nxl))

  def transpose_bwd(self):
    if self.fused_bwd:
      return
    t = _dev.torch()
    P, R0 = self.P, self.ranks[0]
    K, kyl, nxl, nyh = R0.K, R0.kyl, R0.nxl, R0.nyh
    send = [R.ztB.view(K, kyl, P, nxl).permute(2, 0, 1, 3).contiguous() for R in self.ranks]   # [dest][K][kyl][nxl]
    for d, R in enumerate(self.ranks):
      recv = t.stack([send[s][d] for s in range(P)], dim=0)                           # [src = owner of the ky range][K][kyl][nxl]
      R.zt[:, :, 1:nxl + 1] = recv.permute(1, 0, 2, 3).reshape(K, P * kyl, nxl)[:, :nyh]

  def allreduce(self, vecs):
    tot = sum(v.clone() for v in vecs)
    return [tot.clone() for _ in vecs]

  def sync(self):
    pass

  def sums_begin(self, n=None):
    """Totals over the ranks of the first n rows of every rank's `sums` (all of it if n is None): enqueue ..."""
    return sum((R.sums if n is None else R.sums[:n]).clone() for R in self.ranks)

  def sums_end(self, tok):
    """... and read them on the host."""
    return tok.cpu().numpy()

  def allreduce_sums(self, n=None):
    return self.sums_end(self.sums_begin(n))


class DistGroup:
  """One real rank per process; torch.distributed (NCCL) collectives."""

  def __init__(self, rank_state, dist):
    self.ranks = [rank_state]
    self.dist = dist
    self.P = dist.get_world_size()
    self.fused_fwd = self.fused_bwd = False

  def halo(self, get, fenced=False):
    """Ghost rows of every array `get(R)` returns, ONE message per neighbour: the boundary rows of all arrays are packed into
    one send buffer per direction (a grouped send/recv of 2 messages instead of 2 per array)."""
    t = _dev.torch()
    dist, R = self.dist, self.ranks[0]
    P, r = self.P, R.rank
    arrs = list(get(R))
    if P == 1:
      for a in arrs:
        a[..., 0, :] = a[..., R.nxl, :]
        a[..., R.nxl + 1, :] = a[..., 1, :]
      return
    left, right = (r - 1) % P, (r + 1) % P
    firsts = [a[..., 1, :].reshape(-1) for a in arrs]
    lasts = [a[..., R.nxl, :].reshape(-1) for a in arrs]
    s_first, s_last = (firsts[0], lasts[0]) if len(arrs) == 1 else (t.cat(firsts), t.cat(lasts))
    g_left, g_right = t.empty_like(s_last), t.empty_like(s_first)
    # (order matters when left == right, P = 2: NCCL matches the sends and receives of a peer pair in issue order, and the
    #  left ghost has to receive the neighbour's LAST interior row)
    ops = [dist.P2POp(dist.isend, s_last, right), dist.P2POp(dist.isend, s_first, left),
           dist.P2POp(dist.irecv, g_left, left), dist.P2POp(dist.irecv, g_right, right)]
    for w in dist.batch_isend_irecv(ops):
      w.wait()
    off = 0
    for a, f in zip(arrs, firsts):
      n = f.numel()
      a[..., 0, :] = g_left[off:off + n].view(a[..., 0, :].shape)
      a[..., R.nxl + 1, :] = g_right[off:off + n].view(a[..., 0, :].shape)
      off += n

  def _a2a(self, send):
    t = _dev.torch()
    recv = t.empty_like(send)
    self.dist.all_to_all_single(t.view_as_real(recv), t.view_as_real(send))
    return recv

  def transpose_fwd(self):
    t = _dev.torch()
    R, P = self.ranks[0], self.P
    K, kyl, nxl, nyh = R.K, R.kyl, R.nxl, R.nyh
    if getattr(R, "ztpad", None) is None:
      R.ztpad = t.zeros((P, K, kyl, nxl), dtype=R.zt.dtype, device=R.dev)          # [dest][K][kyl][nxl]; rows beyond nyh stay zero
    nfull, rem = nyh // kyl, nyh % kyl
    if nfull:
      R.ztpad[:nfull].copy_(R.zt[:, :nfull * kyl, 1:nxl + 1].reshape(K, nfull, kyl, nxl).permute(1, 0, 2, 3))
    if rem:
      R.ztpad[nfull, :, :rem].copy_(R.zt[:, nfull * kyl:nyh, 1:nxl + 1])
    recv = self._a2a(R.ztpad)
    R.ztB.copy_(recv.permute(1, 2, 0, 3).reshape(K, kyl, P * nxl))

  def transpose_bwd(self):
    R, P = self.ranks[0], self.P
    K, kyl, nxl, nyh = R.K, R.kyl, R.nxl, R.nyh
    recv = self._a2a(R.ztB.view(K, kyl, P, nxl).permute(2, 0, 1, 3).contiguous())
    R.zt[:, :, 1:nxl + 1] = recv.permute(1, 0, 2, 3).reshape(K, P * kyl, nxl)[:, :nyh]

  def allreduce(self, vecs):
    self.dist.all_reduce(vecs[0])
    return vecs

  def sync(self):
    pass

  def sums_begin(self, n=None):
    R = self.ranks[0]
    x = R.sums if n is None else R.sums[:n]
    self.dist.all_reduce(x)
    return x

  def sums_end(self, tok):
    return tok.cpu().numpy()

  def allreduce_sums(self, n=None):
    return self.sums_end(self.sums_begin(n))


class SymmGroup(DistGroup):
  """One rank per process, exchanges over NVLink PEER MEMORY instead of NCCL calls: every exchanged buffer lives in a symmetric-memory
  arena (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into every rank of the node), so
    * a halo exchange   = two strided copy kernels that LOAD the neighbours' boundary rows straight into this rank's ghost rows,
    * an FFT transpose  = P strided copies of this rank's [ky-range of d] x [own x-rows] block straight into rank d's ky-slab (no pack,
                          no permute / contiguous staging, no unpack),
    * the sum all-reduce = P stores of this rank's 84 totals into slot [rank] of every peer + one local sum in fixed rank order (every
                          rank adds the same numbers in the same order: bit-identical totals, hence identical decisions),
  each followed by ONE device-side barrier across the ranks (signal pads, stream-ordered: hdl.barrier()).  Measured on 2 B200s
  (scripts/symm_probe.py): 13 us for a 48 KB peer store + barrier, against ~200 us for the grouped NCCL send/recv of the same rows."""

  def __init__(self, rank_state, dist):
    super().__init__(rank_state, dist)
    import torch.distributed._symmetric_memory as symm_mem
    self.hdl = symm_mem.rendezvous(rank_state.arena, dist.group.WORLD)
    t = _dev.torch()
    n = rank_state.arena.numel()
    self.peer_arena = [self.hdl.get_buffer(r, (n,), t.float64) for r in range(self.P)]
    self.base = rank_state.arena.storage_offset()
    self.set = 0
    self._cache = {}
    R = rank_state
    zb, zr = t.view_as_real(R.ztB), t.view_as_real(R.zt)
    self.fused_fwd, self.fused_bwd = _set_exchange([R], lambda R_, d: self._peer(zb, d).data_ptr(), lambda R_, d: self._peer(zr, d).data_ptr(),
                                                   _fused_default())
    self.side = t.cuda.Stream(device=rank_state.dev)
    self.ev_main, self.ev_side = t.cuda.Event(), t.cuda.Event()

  @staticmethod
  def allocator(device):
    import torch.distributed._symmetric_memory as symm_mem
    t = _dev.torch()
    return lambda n: symm_mem.empty(int(n), dtype=t.float64, device=t.device("cuda", device))

  def _peer(self, a, r):
    """The view of rank r's arena that corresponds to the local arena view `a` (real dtype)."""
    t = _dev.torch()
    p = self.peer_arena[r]
    return t.as_strided(p, a.shape, a.stride(), p.storage_offset() + a.storage_offset() - self.base)

  def _pairs(self, key, build):
    """(destination view in a peer's arena, source view) pairs of one exchange, built once: slicing and as_strided cost the host
    several microseconds per view, and an iteration has ~10 of these copies on its critical path."""
    p = self._cache.get(key)
    if p is None:
      p = self._cache[key] = build()
    return p

  def halo(self, get, fenced=False):
    """PULL model: after a barrier (skipped if the caller says one already separates the neighbours' producing kernels from this
    call: `fenced`), this rank loads the neighbours' boundary rows into its own ghost rows.  (A push would race: the kernels that
    produce an array also write its ghost rows - with values that mean nothing - so a neighbour's early store could be overwritten
    by this rank's own late kernel.  The sources of a pull are interior rows, which only their owner writes, many barriers later.)"""
    R, P = self.ranks[0], self.P
    arrs = get(R)

    def build():
      left, right = (R.rank - 1) % P, (R.rank + 1) % P
      out = []
      for a in arrs:
        out.append((a[..., 0, :], self._peer(a[..., R.nxl, :], left)))            # lower ghost row <- left neighbour's last interior row
        out.append((a[..., R.nxl + 1, :], self._peer(a[..., 1, :], right)))       # upper ghost row <- right neighbour's first interior row
      return out

    if not fenced:
      self.hdl.barrier()
    for dst, src in self._pairs(("halo",) + tuple(a.data_ptr() for a in arrs), build):
      dst.copy_(src)

  def transpose_fwd(self):
    R, P = self.ranks[0], self.P
    if self.fused_fwd:        # phase A stored its spectrum straight into the peers' ky-slabs: only the barrier is left
      self.hdl.barrier()
      return

    def build():
      t = _dev.torch()
      kyl, nxl, nyh = R.kyl, R.nxl, R.nyh
      zr, zb = t.view_as_real(R.zt), t.view_as_real(R.ztB)                # [K][nyh][nxp][2], [K][kyl][nx][2]
      out = []
      for d in range(P):
        k0 = min(d * kyl, nyh)
        kn = max(0, min(kyl, nyh - k0))
        if kn > 0:
          out.append((self._peer(zb[:, :kn, R.rank * nxl:(R.rank + 1) * nxl, :], d), zr[:, k0:k0 + kn, 1:nxl + 1, :]))
      return out

    for dst, src in self._pairs("fwd", build):
      dst.copy_(src)
    self.hdl.barrier()

  def transpose_bwd(self):
    R, P = self.ranks[0], self.P
    if self.fused_bwd:        # phase B stored its result straight into the peers' x-slabs
      self.hdl.barrier()
      return

    def build():
      t = _dev.torch()
      nxl = R.nxl
      zr, zb = t.view_as_real(R.zt), t.view_as_real(R.ztB)
      if R.kyn <= 0:
        return []
      return [(self._peer(zr[:, R.ky0:R.ky0 + R.kyn, 1:nxl + 1, :], d), zb[:, :R.kyn, d * nxl:(d + 1) * nxl, :]) for d in range(P)]

    for dst, src in self._pairs("bwd", build):
      dst.copy_(src)
    self.hdl.barrier()

  def sync(self):
    """All ranks' earlier work on their streams is complete before any rank's later work starts (device-side, no host wait)."""
    self.hdl.barrier()

  def sums_begin(self, n=None):
    """Stores this rank's rows into slot [rank] of every peer, barrier, then copies all P slots to pinned host memory on a SIDE
    stream: the host can wait for exactly that copy (sums_end) while later work is already queued on the main stream."""
    t = _dev.torch()
    R, P = self.ranks[0], self.P
    x = R.sums if n is None else R.sums[:n]
    m = x.numel()
    s = self.set
    self.set ^= 1                                   # two alternating sets: a fast rank's next pass never overwrites what a slow one still sums

    def build():
      mine = R.sums_all[s, R.rank, :m]
      stage = t.empty((P, m), dtype=t.float64, device=R.dev)
      host = t.empty((P, m), dtype=t.float64, pin_memory=True)
      return [(self._peer(mine, r), x.reshape(-1)) for r in range(P)] + [(stage, R.sums_all[s, :, :m]), (host, stage)]

    pairs = self._pairs(("sums", s, m), build)
    for dst, src in pairs[:-2]:
      dst.copy_(src)
    self.hdl.barrier()
    pairs[-2][0].copy_(pairs[-2][1])
    self.ev_main.record()
    with t.cuda.stream(self.side):
      self.side.wait_event(self.ev_main)
      pairs[-1][0].copy_(pairs[-1][1], non_blocking=True)
      self.ev_side.record(self.side)
    return pairs[-1][0], tuple(x.shape)

  def sums_end(self, tok):
    host, shape = tok
    self.ev_side.synchronize()
    rows = host.numpy()                             # [P][m]: every rank adds the same rows in the same (rank) order
    tot = rows[0].copy()
    for r in range(1, self.P):
      tot += rows[r]
    return tot.reshape(shape)

  def allreduce_sums(self, n=None):
    return self.sums_end(self.sums_begin(n))


def make_dist_rank(rank, world, dist, *args, kind=None, **kw):
  """SlabRank + its exchange group for a one-process-per-GPU launch.  kind (default: env PDHG_SLAB_GROUP, else "symm"): "symm" = NVLink
  peer-memory exchanges (SymmGroup), "nccl" = NCCL collectives (DistGroup)."""
  kind = kind or os.environ.get("PDHG_SLAB_GROUP", "symm")
  if kind == "symm":
    R = SlabRank(rank, world, *args, arena_alloc=SymmGroup.allocator(kw.get("device", 0)), **kw)
    return R, SymmGroup(R, dist), kind
  R = SlabRank(rank, world, *args, **kw)
  return R, DistGroup(R, dist), kind


def init_block(group, g_global, c_on_rho):
  """phi0 = tile(g), rho0 = c_on_rho, alp0 = 0 on every rank's slab (utils_pdhg_solver.py:123-137)."""
  t = _dev.torch()
  for R in group.ranks:
    gl = t.from_numpy(np.ascontiguousarray(g_global[R.i0:R.i0 + R.nxl])).to(R.dev)
    for buf in R.phi + [R.phib]:
      buf.zero_()
      buf[:, 1:R.nxl + 1, :] = gl
    for b in R.rho:
      b.fill_(float(c_on_rho))
    for b in R.alp:
      b.zero_()
    R.cp, R.cd = 0, 0
  group.halo(lambda R: R.phi + [R.phib])
  group.sync()


STATS = {}       # of the last solve_block_slab call: outer iterations and how many of them had to redo their dual loop pass by pass
PROFILE = {}     # PDHG_SLAB_PROF=1: seconds per section of solve_block_slab (synchronising timers; diagnostic only)


def _tick(name, t0):
  if t0 is None:
    return None
  _dev.torch().cuda.synchronize()
  now = time.perf_counter()
  PROFILE[name] = PROFILE.get(name, 0.0) + now - t0
  return now


# slots of the sums of sweep s = 1, 2, .. of a fused pass: [rho, alp 1..4] numerators; the denominators follow each
_NUM = np.array([[b0 + 2 * q for q in range(5)] for b0 in [0] + [20 + 16 * sw for sw in range((NQ - 20) // 16)]])
_DEN = _NUM + 1


def _first_hit(v, ns, eps):
  """First sweep (0-based) of a pass of ns fused sweeps whose relative change (update_fns_in_pdhg.py:170-177: rho term + the four
  control terms) is below eps, or -1.  `v`: the NQ totals of the pass.  (Vectorised over the sweeps; the five ratios are added in
  the order of the scalar expression  r0 + (((r1 + r2) + r3) + r4), so the result has the bits of the sweep-by-sweep test.)"""
  with np.errstate(all="ignore"):
    r = v[_NUM[:ns]] / v[_DEN[:ns]]
    e = r[:, 0] + (((r[:, 1] + r[:, 2]) + r[:, 3]) + r[:, 4])
  h = np.flatnonzero(e < eps)
  return int(h[0]) if h.size else -1


def _last_sweep_sums(v, ns):
  """v with slots 0..15 replaced by the sums of sweep ns of a fused pass (a copy when ns > 1)."""
  if ns > 1:
    v = v.copy()
    v[:16] = v[20 + 16 * (ns - 2):36 + 16 * (ns - 2)]
  return v


def solve_block_slab(group, epsl, stepsz_param, n_maxiter, eps=1e-6, rho_alp_iters=10):
  """PDHG_solver_oneiter (utils_pdhg_solver.py:9-94) on the slab-decomposed block.  Returns (iters, end_reason, err1, err2, n_inner)."""
  t = _dev.torch()
  prof = os.environ.get("PDHG_SLAB_PROF") is not None
  tau, sigma = stepsz_param / 1.5, stepsz_param * 1.5
  ranks = group.ranks
  # norms of the starting iterate (interior rows), all-reduced
  loc = []
  for R in ranks:
    it_ = R.interior
    v = t.stack([(it_(R.phi[R.cp])[0] ** 2).sum(), (it_(R.rho[R.cd]) ** 2).sum()] +
                [(it_(R.alp[R.cd][q]) ** 2).sum() for q in range(4)])
    loc.append(v)
  tot = group.allreduce(loc)[0].cpu().numpy()
  group.sync()
  S_row0, S_rho, S_alp = tot[0], tot[1], tot[2:6].copy()
  n_inner, reason, err1, err2 = 0, _lib.END_MAXITER, float("nan"), float("nan")
  prev_j = rho_alp_iters
  spec = os.environ.get("PDHG_SLAB_SPEC", "1") != "0"
  n_respec = 0
  trace = os.environ.get("PDHG_SLAB_TRACE") is not None and ranks[0].rank == 0
  ahead = spec and not prof and os.environ.get("PDHG_SLAB_AHEAD", "1") != "0"
  front_for = None          # dual buffer for which the first half of the next iteration is already in the stream

  def front(b, tk=None):
    """First half of an outer iteration on dual buffer b: ghost rows of (rho, alp1_x, alp2_x), continuity residual + y-transform,
    transpose, x-transform / time solve on this rank's ky-slab, transpose back.  Writes only zt, ztB and ghost rows."""
    # (the barrier of the preceding sum exchange - or the one at the start of the solve - already orders the neighbours' dual
    #  sweeps before this point)
    group.halo(lambda R: [R.dual[b][0:3]], fenced=True)
    tk = _tick("halo1", tk)
    for R in ranks:
      R.ext(R.hL, 0, 0.0, epsl, rho_in=R.rho[b], alp_in=R.alp[b], zt=R.zt)
    tk = _tick("A", tk)
    group.transpose_fwd()
    tk = _tick("a2a_fwd", tk)
    for R in ranks:
      if R.kyn > 0:
        R.ext(R.hB, 1, 0.0, epsl, zt=R.ztB, nyh_override=R.kyn, ky_off=R.ky0, nyh_tab=R.nyh)
    tk = _tick("B", tk)
    group.transpose_bwd()
    return _tick("a2a_bwd", tk)

  it = 0
  for it in range(n_maxiter):
    tk = None
    if prof:
      t.cuda.synchronize()
      tk = time.perf_counter()
    if front_for != ranks[0].cd:
      tk = front(ranks[0].cd, tk)
    front_for = None
    for R in ranks:
      # phase C of the local handle normalises the inverse transforms by 1 / (nxp ny); the x-transform ran over the global nx
      R.ext(R.hL, 2, tau * R.nxp / R.nx, epsl, zt=R.zt, phi_in=R.phi[R.cp], phi_out=R.phi[R.cp ^ 1], phib=R.phib)
    tk = _tick("C", tk)
    group.halo(lambda R: [R.phib])
    tk = _tick("halo2", tk)
    # inner dual loop, as in the single-GPU kernel: buffer cd stays intact, the passes ping-pong between the other two; while
    # the previous outer iteration needed several sweeps, up to `fuse` sweeps are fused per pass (one launch)
    cd = ranks[0].cd
    f1, f2 = (cd + 1) % 3, (cd + 2) % 3
    other = lambda b: f2 if b == f1 else f1

    first_hit = lambda v, ns: _first_hit(v, ns, eps)
    last_sweep_sums = _last_sweep_sums

    def d_pass(src, dst, ns, slot):
      for R in ranks:
        R.ext(R.hL, 3, sigma, epsl, pass_mask=ns, phib=R.phib, rho_in=R.rho[src], alp_in=R.alp[src], rho_out=R.rho[dst], alp_out=R.alp[dst],
              sums=R.sums[slot])

    def e_pass(last, slot):
      for R in ranks:
        R.ext(R.hL, 4, 0.0, epsl, rho_in=R.rho[cd], alp_in=R.alp[cd], rho_out=R.rho[last], alp_out=R.alp[last], sums=R.sums[slot])

    done_spec = False
    if spec:
      # SPECULATIVE pass plan: the sweep count of the previous outer iteration predicts this one's, so ALL its passes and the
      # outer-error pass are enqueued back to back, every launch writing its own row of `sums`, and the ranks exchange the rows
      # ONCE: one all-reduce and one host read per outer iteration instead of one per pass.  The host then replays the exit tests
      # on the rows; if the loop would have stopped anywhere but at the end of the plan, the plan's results are discarded and
      # the loop is redone pass by pass from the intact buffer cd (the sweeps are deterministic: same bits either way).
      plan, j, last = [], 0, cd
      while j < min(prev_j, rho_alp_iters):
        ns = max(1, min(prev_j - j, rho_alp_iters - j, ranks[0].fuse))
        dst = other(last)
        d_pass(last, dst, ns, len(plan))
        plan.append(ns)
        last, j = dst, j + ns
      if j > 1:
        e_pass(last, len(plan))
      tk = _tick("D_pass", tk)
      tok = group.sums_begin(len(plan) + 1)
      if ahead and it + 1 < n_maxiter:
        # the host's read of the sums and its decisions take ~0.1 ms during which the stream would run dry: the first half of
        # the NEXT iteration (it touches nothing this iteration could still need: zt, ztB, ghost rows) is enqueued first
        front(last)
        front_for = last
      V = group.sums_end(tok)
      tk = _tick("allreduce", tk)
      jj, hit_at = 0, -1
      for p, ns in enumerate(plan):
        h = first_hit(V[p], ns)
        if h >= 0:
          hit_at = jj + h + 1
          break
        jj += ns
      if hit_at == j or (hit_at < 0 and j == rho_alp_iters):
        done_spec = True
        v = last_sweep_sums(V[len(plan) - 1], plan[-1])
        e1s0, e1s1, e1nan = V[0][16], V[0][17], V[0][18]
        d_rho, d_alp = v[0], [v[2 + 2 * q] for q in range(4)]
        if j > 1:
          vo = V[len(plan)]
          d_rho, d_alp = vo[10], [vo[11 + q] for q in range(4)]
      else:
        n_respec += 1
        front_for = None
        if n_respec > 8 and 4 * n_respec > it:       # the sweep count keeps changing: the prediction does not pay on this problem
          spec = False
    if not done_spec:
      j, last = 0, cd
      while j < rho_alp_iters:
        src, dst = last, other(last)
        ns = max(1, min(prev_j - j, rho_alp_iters - j, ranks[0].fuse))
        d_pass(src, dst, ns, 0)
        tk = _tick("D_pass", tk)
        v = group.allreduce_sums(1)[0]
        tk = _tick("allreduce", tk)
        if j == 0:
          e1s0, e1s1, e1nan = v[16], v[17], v[18]
        hit = first_hit(v, ns)
        done = ns
        if 0 <= hit < ns - 1:
          # the exit falls inside a fused pass: redo exactly the sweeps up to it from the pass's input
          done = hit + 1
          d_pass(src, dst, done, 0)
          v = group.allreduce_sums(1)[0]
        v = last_sweep_sums(v, done)
        last, j = dst, j + done
        if hit >= 0:
          break
      d_rho, d_alp = v[0], [v[2 + 2 * q] for q in range(4)]
      if j > 1:
        e_pass(last, 0)
        vo = group.allreduce_sums(1)[0]
        d_rho, d_alp = vo[10], [vo[11 + q] for q in range(4)]
    prev_j = j
    n_inner += j
    with np.errstate(all="ignore"):
      err1 = np.sqrt(e1s0) / np.sqrt(S_row0 + e1s1)
      err2 = np.sqrt(d_rho) / np.sqrt(S_rho)
      for q in range(4):
        na, ne = np.sqrt(S_alp[q]), np.sqrt(d_alp[q])
        if na < 1e-6 and ne > 1e-6:
          err2 += ne
        elif na >= 1e-6:
          err2 += ne / na
    if trace:
      print("slab it %d err1 %.17g err2 %.17g sweeps %d spec %s" % (it, err1, err2, j, done_spec), flush=True)
    S_rho, S_alp = v[1], np.array([v[3 + 2 * q] for q in range(4)])
    for R in ranks:
      R.cp ^= 1
      R.cd = last
    if err1 < eps and err2 < eps:
      reason = _lib.END_CONVERGED
      break
    if e1nan > 0 or v[15] > 0:
      reason = _lib.END_NAN
      break
  iters = it + 1
  STATS["iters"], STATS["respeculated"] = iters, n_respec
  return iters, reason, float(err1), float(err2), n_inner


def gather_block(group):
  """Assembles phi [2,nx,ny], rho [1,nx,ny], alp [4,1,nx,ny,2] of the block on the host (LocalGroup) / this rank's slab (DistGroup)."""
  t = _dev.torch()
  phi = t.cat([R.interior(R.phi[R.cp]) for R in group.ranks], dim=1).cpu().numpy()
  rho = t.cat([R.interior(R.rho[R.cd]) for R in group.ranks], dim=1).cpu().numpy()
  al = t.cat([R.interior(R.alp[R.cd]) for R in group.ranks], dim=2).cpu().numpy()        # [4,1,nx(l),ny] active components
  alp = np.zeros(al.shape + (2,))
  alp[:2, ..., 0] = al[:2]
  alp[2:, ..., 1] = al[2:]
  return phi, rho, alp

"""bench.py — headline benchmark of the B200-native PDHG hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" = ONE outer PDHG iteration (primal phi update through the spectral preconditioner, extrapolation, dual
rho/alp sweeps, error reductions and exit tests; jaxsrc/utils/utils_pdhg_solver.py:51-88) on the named workload.
Headline workload (BASELINE.json configs[2], the config the north-star's HBM-roofline target is quoted on):
  cfg3_tsp65 = egno=1 ndim=2 epsl=0 nx=ny=256 nt=65 with --time_step_per_PDHG 65 (one space-time block,
  N = 64*256*256 = 4.19 M points, ~300 MB of state: larger than the 126 MB L2, i.e. HBM-bound).
Other workloads (secondary lines in the "others" key or via --workload): cfg3_tsp2 (reference default tsp=2,
L2-resident), cfg1 (README example, full solve), cfg2 (1-D viscous 640x161 at a stable step size).
N > 1 (torchrun): the sharded workload of the north-star — BASELINE configs[3], 4096 independent 1-D instances, first 8 time
blocks, contiguous instance ranges per rank, NO data-path collective (strong scaling; logs all-gathered over NCCL) — plus, under
"others", the x-slab decomposition of the configs[4] grid over the same ranks and the same instances on one GPU.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pdhg-optimal-control_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)

WORKLOADS = {
  # name: (egno, ndim, nx, ny, nt, tsp, epsl, stepsz)
  "cfg3_tsp65": (1, 2, 256, 256, 65, 65, 0.0, 0.05),
  "cfg3_tsp2": (1, 2, 256, 256, 65, 2, 0.0, 0.1),
  "cfg2": (1, 1, 640, 1, 161, 2, 0.1, 0.005),
  "cfg1": (1, 1, 160, 1, 41, 2, 0.0, 0.1),
  "cfg5_tsp2": (1, 2, 2048, 2048, 257, 2, 0.1, 5e-4),
}
DESCR = {
  "cfg3_tsp65": "BASELINE configs[2]: egno=1 ndim=2 epsl=0 nx=ny=256 nt=65, time_step_per_PDHG=65 (one space-time block, HBM-bound)",
  "cfg3_tsp2": "BASELINE configs[2]: egno=1 ndim=2 epsl=0 nx=ny=256 nt=65, time_step_per_PDHG=2 (reference default, L2-resident block 0)",
  "cfg2": "BASELINE configs[1]: egno=1 ndim=1 epsl=0.1 nx=640 nt=161 tsp=2 at the stable stepsz_param=0.005 (block 0)",
  "cfg1": "BASELINE configs[0]: egno=1 ndim=1 epsl=0 nx=160 nt=41 tsp=2 stepsz_param=0.1 (block 0)",
  "cfg5_tsp2": "BASELINE configs[4] grid on ONE GPU: egno=1 ndim=2 epsl=0.1 nx=ny=2048 nt=257 tsp=2, stepsz_param=5e-4 (block 0, fixed "
               "iteration budget; generic 2048-point transforms; the x-slab decomposition is scripts/slab_bench.py)",
}


def algorithmic_bytes_per_iter(ndim, K, nx, ny, n_in=1.0):
  """8*N*[(1+A)+1 + 2+2 + n_in*((2+1+A)+(1+A))]  (fp64, active control components only)."""
  A = 2 * ndim
  N = K * nx * ny
  return 8.0 * N * ((1 + A) + 1 + 4 + n_in * ((3 + A) + (1 + A)))


def measured_peak_gbs():
  f = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(f):
    try:
      return float(json.load(open(f))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
      pass
  return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
  Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

  def __init__(self, index):
    self.index, self.rows, self.proc, self.th = index, [], None, None

  def _run(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                    "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      for line in self.proc.stdout:
        self.rows.append([c.strip() for c in line.split(",")])
    except Exception:
      pass

  def start(self):
    self.th = threading.Thread(target=self._run, daemon=True)
    self.th.start()
    time.sleep(0.25)

  def stop(self):
    time.sleep(0.15)
    try:
      if self.proc:
        self.proc.terminate()
    except Exception:
      pass
    if self.th:
      self.th.join(timeout=2)
    sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k].lower().startswith("active") for r in self.rows)]
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
            "samples": len(sm)}


def make_problem(name):
  from pdhg_b200 import run_example as rx, set_fns
  egno, ndim, nx, ny, nt, tsp, epsl, stepsz = WORKLOADS[name]
  n_ctrl, bc, cen = rx.problem_setup(egno, ndim)
  x_arr = rx.make_x_arr(ndim, nx, ny, 2.0, 2.0, cen)
  with contextlib.redirect_stdout(io.StringIO()):
    fns = set_fns.set_up_example_fns(egno, ndim, 0)
  dt, period, dspatial, nspatial = rx._grid(ndim, nx, ny, nt, 2.0, 2.0, 1.0)
  g = set_fns.set_up_J(egno, ndim, period)(x_arr)
  return dict(name=name, egno=egno, ndim=ndim, nx=nx, ny=ny, nt=nt, tsp=tsp, K=tsp - 1, epsl=epsl, stepsz=stepsz, n_ctrl=n_ctrl, bc=bc,
              x_arr=x_arr, fns=fns, dt=dt, dspatial=dspatial, nspatial=nspatial, g=g, N=(tsp - 1) * nx * ny)


def run_ours_block(pb, steps, warmup, device, epsl_shift=0.0, sampler=None, barrier=None, spinup=0):
  """Times `steps` consecutive outer iterations of block 0 with the state resident in HBM (kernel launched through the
  C ABI): iterations [spinup + warmup, spinup + warmup + steps) of the solve that starts from the cold initial state
  (phi0 = tile(g), rho0 = c_on_rho, alp0 = 0); the `spinup + warmup` iterations before them run untimed."""
  import torch
  from pdhg_b200.update_fns_in_pdhg import get_solver
  torch.cuda.set_device(device)
  s = get_solver(pb["fns"], pb["nspatial"], pb["K"], pb["bc"], pb["dt"], pb["dspatial"], 70.0, pb["x_arr"], nblocks=1, max_rec=4, device=device)
  K, nsp, A, n_ctrl = pb["K"], tuple(pb["nspatial"]), 2 * pb["ndim"], pb["n_ctrl"]
  dev = torch.device("cuda", device)
  g = torch.from_numpy(np.ascontiguousarray(pb["g"])).to(dev)
  phi0 = g.expand((K + 1,) + nsp).contiguous()
  rho0 = torch.full((K,) + nsp, 70.0, dtype=torch.float64, device=dev)
  alp0 = torch.zeros((A, K) + nsp + (n_ctrl,), dtype=torch.float64, device=dev)
  po, ro, ao = torch.empty_like(phi0), torch.empty_like(rho0), torch.empty_like(alp0)
  epsl = pb["epsl"] + epsl_shift
  stream = torch.cuda.current_stream(dev).cuda_stream

  def call(n_it, begin=0, src=None):
    p0, r0, a0 = src if src is not None else (phi0, rho0, alp0)
    return s.solve_block_dev(p0.data_ptr(), r0.data_ptr(), a0.data_ptr(), epsl, pb["stepsz"], begin + n_it, begin, 0, 0,
                             po.data_ptr(), ro.data_ptr(), ao.data_ptr(), stream)
  # untimed: `spinup` + `warmup` iterations from the cold initial state; the timed iterations continue from there
  begin = spinup + max(warmup, 3)
  call(begin)
  src = (po.clone(), ro.clone(), ao.clone())
  torch.cuda.synchronize(dev)
  if barrier:
    barrier()
  if sampler:
    sampler.start()
  l0 = s.launch_count
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize(dev)
  e0.record()
  logs = call(steps, begin, src)
  e1.record()
  torch.cuda.synchronize(dev)
  if barrier:
    barrier()
  clocks = sampler.stop() if sampler else None
  ms = e0.elapsed_time(e1)
  iters = int(logs.iters[0, 0]) - begin
  return dict(begin=begin, ms=ms, kernel_ms=s.last_kernel_ms, iters=iters, n_inner=int(logs.inner_total[0]), launches=s.launch_count - l0,
              end_reason=int(logs.end_reason[0, 0]), clocks=clocks, path=s.path, solver=s,
              state=src)


def run_ours_e2e(pb, steps, device, state):
  """Same iterations end to end through the reference-facing call with HOST buffers:
  PDHG_solver_oneiter(fn_update_primal, fn_update_dual, fns_dict, phi0, rho0, alp0, ...) with NumPy arrays in (pinned host
  memory) and NumPy arrays out; H2D of the state, `steps` iterations, D2H of the result are all inside the timed region."""
  import torch
  from pdhg_b200.update_fns_in_pdhg import NativeUpdateDual, NativeUpdatePrimal
  from pdhg_b200.utils.utils_pdhg_solver import PDHG_solver_oneiter
  pin = lambda t: t.cpu().pin_memory().numpy()
  phi0, rho0 = pin(state[0]), pin(state[1])
  alp0 = tuple(pin(state[2][j]) for j in range(state[2].shape[0]))
  P, D = NativeUpdatePrimal(pb["ndim"], pb["bc"]), NativeUpdateDual(pb["bc"])
  args = (P, D, pb["fns"], phi0, rho0, alp0, pb["x_arr"], None, pb["ndim"], pb["dt"], pb["dspatial"], 70.0)
  kw = dict(epsl=pb["epsl"], stepsz_param=pb["stepsz"], fv=None, N_maxiter=steps, print_freq=0, eps=1e-6)
  with contextlib.redirect_stdout(io.StringIO()):
    PDHG_solver_oneiter(*args, **dict(kw, N_maxiter=3))      # warm the handle / allocations
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res, errs = PDHG_solver_oneiter(*args, **kw)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
  it, phi, rho, alp = res[-1]
  h2d = phi0.nbytes + rho0.nbytes + sum(a.nbytes for a in alp0)
  d2h = phi.nbytes + rho.nbytes + sum(a.nbytes for a in alp)
  return dict(s=t, iters=int(it), h2d=h2d, d2h=d2h)


def oracle_iterations(pb, n_iters):
  """Times `n_iters` outer iterations of the NumPy oracle (the CPU restatement of the reference) on block 0."""
  from oracle import pdhg_numpy as orc
  egno, ndim, K = pb["egno"], pb["ndim"], pb["K"]
  fns = orc.set_up_example_fns(egno, ndim, 0)
  bc, nsp, dsp = pb["bc"], pb["nspatial"], pb["dspatial"]
  fv = orc.compute_Dxx_fft_fv(ndim, nsp, dsp, bc)
  up = orc.update_primal_1d if ndim == 1 else orc.update_primal_2d
  prim = lambda *a: up(*a, bc)
  dual = lambda *a, eps: orc.update_dual_alternative(*a, bc, eps=eps)
  phi0 = np.concatenate([pb["g"]] * (K + 1), axis=0)
  rho0 = np.zeros((K,) + tuple(nsp)) + 70.0
  alp0 = tuple(np.zeros((K,) + tuple(nsp) + (pb["n_ctrl"],)) for _ in range(2 * ndim))
  t0 = time.perf_counter()
  res, _ = orc.PDHG_solver_oneiter(prim, dual, fns, phi0, rho0, alp0, pb["x_arr"], None, ndim, pb["dt"], dsp, 70.0, epsl=pb["epsl"],
                                   stepsz_param=pb["stepsz"], fv=fv, N_maxiter=n_iters, print_freq=0, eps=1e-6)
  return time.perf_counter() - t0, res[-1][0]


def cpu_baseline(pb, budget_s=20.0):
  """NumPy oracle (the CPU port of the reference) on the SAME iterations as the main line (the first outer iterations of block 0
  from the cold initial state), at least 3 of them, 1 process: NumPy's elementwise kernels and pocketfft are single-threaded."""
  t1, it1 = oracle_iterations(pb, 1)
  n = int(max(3, min(200, budget_s / max(t1, 1e-6))))
  t, it = oracle_iterations(pb, n)
  return {"value": it * pb["N"] / t, "unit": "grid-point updates/s", "cores": 1, "kind": "port",
          "iters_per_s": it / t, "sample": "the first %d outer PDHG iterations of block 0 of the same workload from the cold initial state (the main line's "
          "window), NumPy fp64 oracle, 1 process (JAX is not installable here, so the reference's own JAX-CPU path cannot be timed)" % it}


def cfg4_inputs(B, nx=1024):
  rng = np.random.default_rng(0)
  A_, th, u = rng.uniform(0.5, 1.5, 4096)[:B], rng.uniform(0, 2 * np.pi, 4096)[:B], rng.uniform(0, 1, 4096)[:B]
  x = np.linspace(0.0, 2.0, num=nx, endpoint=False)
  g = A_[:, None] * np.sin(np.pi * x[None, :] + th[:, None])
  return g, 0.002 * u


def _cfg4_cpu_worker(args):
  """One oracle process: marches instance `b` of configs[3] through `nblk` time blocks; returns (iterations, seconds)."""
  b, nblk, nx, nt_full = args
  from oracle import pdhg_numpy as orc
  g, epsl = cfg4_inputs(b + 1, nx)
  x_arr, bc, n_ctrl = orc.make_grid(1, 1, nx, 1, 2.0, 2.0)
  info = {}
  t0 = time.perf_counter()
  orc.solve_HJ(1, n_ctrl, 1, float(epsl[b]), orc.set_up_example_fns(1, 1, 0), nx, 1, nblk + 1, 2.0, 2.0, nblk / (nt_full - 1.0), x_arr, 70.0, 2, 0.1,
               1000000, 10 ** 9, 1e-6, bc, g=g[b:b + 1], info=info)
  return int(sum(info["block_iters"])), time.perf_counter() - t0


def cpu_baseline_batched(nblk=1, nx=1024, nt_full=257):
  """Leg (ii) of SURVEY.md section 8(d): os.cpu_count() oracle processes, each marching ONE configs[3] instance through its first
  time block (independent instances need no communication, so this is what the CPU can do with all its cores)."""
  import multiprocessing as mp
  cores = os.cpu_count() or 1
  ctx = mp.get_context("spawn")
  t0 = time.perf_counter()
  with ctx.Pool(cores) as pool:
    res = pool.map(_cfg4_cpu_worker, [(b, nblk, nx, nt_full) for b in range(cores)])
  wall = time.perf_counter() - t0
  its = sum(r[0] for r in res)
  busy = max(r[1] for r in res)
  return {"value": its * nx / busy, "unit": "grid-point updates/s", "cores": cores, "kind": "port", "pdhg_iters_per_s": its / busy, "busy_s": busy,
          "sample": "%d processes (os.cpu_count()), one BASELINE configs[3] instance each (nx=%d), first %d of %d time blocks to the reference's "
                    "stopping rule: %d iterations in %.1f s (slowest process; %.1f s incl. process start-up)" % (cores, nx, nblk, nt_full - 1, its, busy, wall)}


SPINUP = {"cfg3_tsp65": 600}    # untimed iterations before the timed region (steady-state inner-sweep count)


def time_to_tol(name, device, nblocks=None):
  """Wall time of the reference-facing solve_HJ call (host arrays in/out) to the reference's own stopping rule."""
  from pdhg_b200 import run_example as rx
  pb = make_problem(name)
  nt = pb["nt"] if nblocks is None else nblocks * pb["K"] + 1
  T = 1.0 if nblocks is None else pb["dt"] * (nt - 1)
  info = {}
  run = lambda: rx.solve_HJ(pb["ndim"], pb["n_ctrl"], pb["egno"], pb["epsl"], pb["fns"], pb["nx"], pb["ny"], nt, 2.0, 2.0, T, pb["x_arr"], 70.0,
                            pb["tsp"], 0.1, 1000000, 10000, 1e-6, pb["bc"], info=info)
  with contextlib.redirect_stdout(io.StringIO()):
    t0 = time.perf_counter()
    run()
    t = time.perf_counter() - t0
  its = int(sum(info["block_iters"]))
  return {"workload": DESCR[name].split(" (")[0] + (" — all %d time blocks" % (nt - 1) if nblocks is None else " — first %d time blocks" % nblocks)
                      + ", solve_HJ with stepsz_param=0.1 incl. the NaN fallback", "seconds_to_tol": t, "total_iters": its,
          "blocks": len(info["block_iters"]), "stepsz_used": sorted(set(info["stepsz_used"]), reverse=True), "pdhg_iters_per_s": its / t}


def secondary(name, device, iters):
  pb = make_problem(name)
  r = run_ours_block(pb, iters, 3, device, spinup=SPINUP.get(name, 0))
  by = algorithmic_bytes_per_iter(pb["ndim"], pb["K"], pb["nx"], pb["ny"], r["n_inner"] / max(r["iters"], 1)) * r["iters"]
  return {"workload": DESCR[name], "iters": r["iters"], "pdhg_iters_per_s": r["iters"] / (r["ms"] * 1e-3),
          "grid_point_updates_per_s": r["iters"] * pb["N"] / (r["ms"] * 1e-3), "us_per_iter": r["ms"] * 1e3 / max(r["iters"], 1),
          "algorithmic_GBps": by / (r["kernel_ms"] * 1e-3) / 1e9, "kernel_path": r["path"]}


def src_hash():
  """sha256 (first 16 hex digits) of the kernel sources: stamps ncu-derived numbers so that stale ones are never reported."""
  import hashlib
  h = hashlib.sha256()
  d = os.path.join(ROOT, "pdhg-optimal-control_b200", "csrc")
  for fn in sorted(os.listdir(d)):
    if fn.endswith((".cu", ".cuh", ".h")):
      h.update(open(os.path.join(d, fn), "rb").read())
  return h.hexdigest()[:16]


def measured_traffic(tag):
  """DRAM bytes per outer iteration from the committed ncu capture profiles/traffic_<tag>.json — only when that capture was
  taken with exactly these kernel sources (src_hash); otherwise None (never a stale constant)."""
  f = os.path.join(ROOT, "profiles", "traffic_%s.json" % tag)
  try:
    d = json.load(open(f))
    if d.get("src_hash") == src_hash():
      return float(d["dram_bytes_per_iter"])
  except Exception:
    pass
  return None


def roofline_of(pb, r, peak, peak_src, tag):
  n_in = r["n_inner"] / max(r["iters"], 1)
  by = algorithmic_bytes_per_iter(pb["ndim"], pb["K"], pb["nx"], pb["ny"], n_in) * r["iters"]
  ach = by / (r["kernel_ms"] * 1e-3) / 1e9
  tr = measured_traffic(tag)
  out = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
         "traffic": tr * r["iters"] if tr else None,
         "dram_frac": (tr * r["iters"] / (r["kernel_ms"] * 1e-3) / 1e9 / peak) if tr else None,
         "kernel": "pdhg_coop_kernel" if r["path"] == 2 else "pdhg1d_cta_kernel", "kernel_ms": r["kernel_ms"], "launches": r["launches"],
         "iters": r["iters"], "inner_sweeps_per_iter": n_in, "ms_per_iter": r["kernel_ms"] / max(r["iters"], 1),
         "algorithmic_bytes_per_launch": by, "peak_source": peak_src,
         "traffic_source": ("ncu dram__bytes_read+write of this window, profiles/traffic_%s.json (same kernel sources: src_hash %s)" % (tag, src_hash()))
                           if tr else "no ncu capture of these kernel sources committed (profiles/traffic_%s.json is stale or absent)" % tag}
  return out


def cfg4_sharded(rank, world, local, dist, nblk=8, B=4096):
  """BASELINE configs[3]: 4096 independent 1-D instances (nx=1024, varied initial data / epsl), first `nblk` of the 256 time blocks,
  sharded over the ranks with NO data-path collective (pdhg_b200/sharding.py -> solve_HJ_batch per rank, one CTA per instance);
  the per-instance logs are all-gathered at the end.  Strong scaling: the same B instances at every N."""
  import torch
  from pdhg_b200 import run_example as rx, set_fns, sharding
  nx, nt_full = 1024, 257
  g, epsl = cfg4_inputs(B, nx)
  x_arr = rx.make_x_arr(1, nx, 1, 2.0, 2.0)
  with contextlib.redirect_stdout(io.StringIO()):
    fns = set_fns.set_up_example_fns(1, 1, 0)
  T = nblk / (nt_full - 1.0)
  kms = {}

  def solve(gs, es, ss):
    info = {}
    out = rx.solve_HJ_batch(1, 1, 1, es, fns, nx, 1, nblk + 1, 2.0, 2.0, T, x_arr, gs, 70.0, 2, ss, 1000000, 10 ** 9, 1e-6, 0, device=local,
                            info=info)
    kms["ms"] = info["kernel_ms"]
    return out
  run = lambda: sharding.solve_batch_sharded(solve, g, epsl, 0.1, rank, world, dist if world > 1 else None)
  # warm-up on a small slice (handle creation, allocations), then the timed sharded solve
  sharding.solve_batch_sharded(solve, g[:max(world * 4, 8)], epsl[:max(world * 4, 8)], 0.1, rank, world, dist if world > 1 else None)
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  b, e, phi, rho, alp, logs = run()
  torch.cuda.synchronize()
  t_local = time.perf_counter() - t0
  t = torch.tensor([t_local, kms["ms"] * 1e-3], dtype=torch.float64, device="cuda")
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  its = float(np.sum(logs["total_iters"]))
  return {"workload": "BASELINE configs[3]: %d independent 1-D instances (egno=1, nx=1024, nt=257, varied initial data / epsl), first %d of 256 time "
                      "blocks each to the reference's stopping rule, sharded over %d GPU(s) with no data-path collective" % (B, nblk, world),
          "instances": B, "blocks": nblk, "total_iters": its, "all_converged": bool(np.all(logs["status"] == 0)),
          "seconds_e2e": float(t[0]), "seconds_kernel": float(t[1]), "pdhg_iters_per_s": its / float(t[0]),
          "value_device": its * nx / float(t[1]), "value_e2e": its * nx / float(t[0]),
          "h2d_bytes": int(g[b:e].nbytes), "d2h_bytes": int(phi.nbytes + rho.nbytes + alp.nbytes), "my_instances": [int(b), int(e)]}


def slab_line(rank, world, local, dist, nx=2048, iters=30):
  """BASELINE configs[4] grid (2048 x 2048, epsl = 0.1, tsp = 2, stepsz 5e-4, block 0) x-slab decomposed over the ranks (NCCL halo
  exchange, all-to-all transposes, all-reduce of the error sums; pdhg_b200/slab.py): iterations/s of a fixed iteration budget."""
  import torch
  from pdhg_b200 import run_example as rx, set_fns as sf, slab
  ny, T, epsl, stepsz = nx, 1.0 / 256, 0.1, 5e-4
  x_arr = rx.make_x_arr(2, nx, ny, 2.0, 2.0)
  with contextlib.redirect_stdout(io.StringIO()):
    fns = sf.set_up_example_fns(1, 2, 0)
  g = sf.set_up_J(1, 2, (2.0, 2.0))(x_arr)[0]
  try:
    R, grp, kind = slab.make_dist_rank(rank, world, dist, fns, nx, ny, T, (2.0 / nx, 2.0 / ny), 70.0, x_arr, device=local)
  except Exception as ex:          # no symmetric (peer-mapped) memory on this box: the NCCL flavour of the same exchanges, and say so
    sys.stderr.write("slab_line: peer-memory exchanges unavailable (%r), using NCCL collectives\n" % (ex,))
    R, grp, kind = slab.make_dist_rank(rank, world, dist, fns, nx, ny, T, (2.0 / nx, 2.0 / ny), 70.0, x_arr, device=local, kind="nccl")
  out = {}
  for label, n in (("warm", 3), ("timed", iters)):
    slab.init_block(grp, g, 70.0)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = slab.solve_block_slab(grp, epsl, stepsz, n)
    torch.cuda.synchronize(); dist.barrier()
    out[label] = (time.perf_counter() - t0, res)
  t = torch.tensor([out["timed"][0]], dtype=torch.float64, device="cuda")
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  it_s, n_in = out["timed"][1][0], out["timed"][1][4]
  one = None
  if rank == 0:
    # the same block and iteration budget through the single-GPU cooperative kernel, in this very run (device time of the march)
    n_ctrl, bc, _ = rx.problem_setup(1, 2)
    with contextlib.redirect_stdout(io.StringIO()):
      for _ in range(2):
        info = {}
        rx.solve_HJ(2, n_ctrl, 1, epsl, fns, nx, ny, 2, 2.0, 2.0, T, x_arr, 70.0, 2, stepsz, iters, 10 ** 9, 1e-6, bc, info=info)
    if (info.get("kernel_ms") or 0) > 0:
      one = info["kernel_ms"] / max(info["block_iters"][0], 1)
  ms_it = float(t[0]) / max(it_s, 1) * 1e3
  return {"one_gpu_ms_per_iter": one, "speedup_vs_one_gpu": (one / ms_it) if one else None,"workload": "BASELINE configs[4] grid: egno=1 ndim=2 epsl=0.1 nx=ny=%d tsp=2 stepsz_param=5e-4, block 0, x-slab decomposed over %d GPUs "
                      "(2 halo exchanges, 2 transposes of the half spectrum, 1 sum all-reduce per dual pass and iteration)" % (nx, world),
          "exchange": {"symm": "NVLink peer-memory stores + device-side barrier (torch symmetric memory)", "nccl": "NCCL collectives"}[kind],
          "iters": it_s, "inner_sweeps_per_iter": n_in / max(it_s, 1), "seconds": float(t[0]), "pdhg_iters_per_s": it_s / float(t[0]),
          "grid_point_updates_per_s": it_s * nx * ny / float(t[0]), "ms_per_iter": float(t[0]) / max(it_s, 1) * 1e3}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="ours")
  ap.add_argument("--workload", default="cfg3_tsp65")
  ap.add_argument("--no-others", action="store_true")
  a = ap.parse_args()
  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  warm = max(a.warmup, 3)

  if a.impl == "reference":
    if rank != 0:
      return 0
    if world > 1:
      # the N-rank line of our arm is the sharded configs[3] workload: the CPU can run it on all its cores (independent instances)
      cb = cpu_baseline_batched(nblk=1)
      line = {"impl": "reference", "metric": "grid_point_updates_per_s", "value": cb["value"], "unit": "grid-point updates/s", "n_gpus": a.gpus,
              "steps": 1, "requested_steps": a.steps, "warmup": 0, "ms_per_step": cb["busy_s"] * 1e3, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "f64", "data": "synthetic", "pdhg_iters_per_s": cb["pdhg_iters_per_s"],
              "config": {"workload": "BASELINE configs[3]: independent 1-D instances (egno=1, nx=1024, nt=257, varied initial data / epsl) to the "
                                     "reference's stopping rule; CPU sample: one instance per core, first time block"},
              "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "grid-point updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
      print(json.dumps(line))
      return 0
    pb = make_problem_cpu(a.workload)
    t1, _ = oracle_iterations(pb, 1)          # warm-up (also sizes the sample)
    # the steps this arm can finish within ~150 s; `steps` in the line is what was actually timed
    n = int(max(1, min(a.steps, 150.0 / max(t1, 1e-6))))
    t, it = oracle_iterations(pb, n)
    val = it * pb["N"] / t
    line = {"impl": "reference", "metric": "grid_point_updates_per_s", "value": val, "unit": "grid-point updates/s", "n_gpus": a.gpus,
            "steps": it, "requested_steps": a.steps, "warmup": 1, "ms_per_step": t / it * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "pdhg_iters_per_s": it / t,
            "config": {"workload": DESCR[a.workload]},
            "cpu_baseline": {"value": val, "unit": "grid-point updates/s", "cores": 1, "kind": "port",
                             "sample": "the first %d outer PDHG iterations of block 0 from the cold initial state (the window our arm times; %d were "
                                       "requested, %d fit the time budget), NumPy fp64 oracle port of the reference, 1 process; the reference's JAX path "
                                       "cannot be installed in this image" % (it, a.steps, it)},
            "e2e": {"value": val, "unit": "grid-point updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0

  import torch
  import torch.distributed as dist
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device; the product path has no CPU fallback")
  peak, peak_src = measured_peak_gbs()

  if world > 1:
    # ---- N ranks: the sharded workload of the north-star (configs[3], independent instances, no collective); strong scaling ----
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
      sampler.start()
    sh = cfg4_sharded(rank, world, local, dist)
    clocks = sampler.stop() if sampler else None
    others = {}
    if not a.no_others:
      try:
        others["cfg5_slab"] = slab_line(rank, world, local, dist)
      except Exception as ex:
        others["cfg5_slab"] = {"error": repr(ex)}
      dist.barrier()
      if rank == 0:
        try:      # the same instances on ONE GPU in this very run: the base of the strong-scaling curve
          others["cfg4_one_gpu"] = cfg4_sharded(0, 1, local, None)
        except Exception as ex:
          others["cfg4_one_gpu"] = {"error": repr(ex)}
      dist.barrier()
    if rank == 0:
      line = {"metric": "grid_point_updates_per_s", "value": sh["value_device"], "unit": "grid-point updates/s", "n_gpus": world, "steps": a.steps,
              "warmup": a.warmup, "ms_per_step": sh["seconds_kernel"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "f64", "data": "synthetic", "pdhg_iters_per_s": sh["total_iters"] / sh["seconds_kernel"],
              "config": {"workload": sh["workload"], "parallelism": "instances sharded x%d (contiguous ranges, no data-path collective; logs all-gathered)" % world,
                         "step": "ONE step = the whole sharded solve (%d instances x %d time blocks = %.3g PDHG iterations); --steps / --warmup do not "
                                 "apply: the workload runs to the reference's own stopping rule" % (sh["instances"], sh["blocks"], sh["total_iters"]),
                         "l2": "state lives in shared memory / registers (one CTA per instance): HBM is touched once per time block",
                         "kernel_path": "single-CTA K=1 register-resident kernel (pdhg1d_k1_kernel)"},
              "gpu_launches": 2 * world, "clocks": clocks, "all_converged": sh["all_converged"],
              "e2e": {"value": sh["value_e2e"], "unit": "grid-point updates/s", "h2d_bytes_per_step": sh["h2d_bytes"] * world,
                      "d2h_bytes_per_step": sh["d2h_bytes"] * world, "seconds": sh["seconds_e2e"],
                      "call": "sharding.solve_batch_sharded -> run_example.solve_HJ_batch with host arrays in / out on every rank (H2D of g, march, D2H of "
                              "phi / rho / alp of all blocks), barrier + max over ranks"},
              "roofline": {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                           "note": "latency-bound regime by construction (state never leaves the SM, SURVEY.md section 8d); the HBM roofline line is the "
                                   "N = 1 run on BASELINE configs[2]"},
              "others": others}
      print(json.dumps(line))
    dist.destroy_process_group()
    return 0

  # ---- one GPU: BASELINE configs[2] with time_step_per_PDHG = 65 ----
  pb = make_problem(a.workload)
  sampler = ClockSampler(local)
  # main window = iterations [W, W + K) of the solve from its cold initial state — the same iterations the reference arm and the
  # cpu_baseline time on the CPU (they cannot afford the ~600 iterations to the steady state), so value / e2e compare like with like
  r = run_ours_block(pb, a.steps, warm, local, sampler=sampler, spinup=0)
  ms, iters = r["ms"], r["iters"]
  e2e = run_ours_e2e(pb, a.steps, local, r["state"])
  N = pb["N"]
  value = iters * N / (ms * 1e-3)
  line = {"metric": "grid_point_updates_per_s", "value": value, "unit": "grid-point updates/s", "n_gpus": 1, "steps": a.steps,
          "warmup": a.warmup, "ms_per_step": ms / max(iters, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f64", "data": "synthetic", "pdhg_iters_per_s": iters / (ms * 1e-3), "iters_timed": iters,
          "inner_sweeps_per_iter": r["n_inner"] / max(iters, 1), "end_reason": r["end_reason"],
          "config": {"workload": DESCR[a.workload], "grid_points_per_iter": N, "stepsz_param": pb["stepsz"],
                     "timed_region": "iterations %d..%d of block 0 from the cold initial state (up to 10 inner dual sweeps per iteration there: "
                                     "inner_sweeps_per_iter); the roofline object is the STEADY state of the same solve (1 sweep per iteration, > 95 %% of a "
                                     "real solve), measured in this run" % (r["begin"], r["begin"] + iters),
                     "parallelism": "one GPU",
                     "l2": "state ~%.0f MB > 126 MB L2 (inputs larger than L2, no flush needed)" % (N * 8 * 14 / 1e6)
                           if N * 8 * 14 > 126e6 else "state fits L2 (L2-resident regime; no flush: the iteration re-reads its own state)",
                     "kernel_path": "cooperative multi-CTA" if r["path"] == 2 else "single-CTA smem"},
          "gpu_launches": r["launches"], "clocks": r["clocks"]}
  line["e2e"] = {"value": e2e["iters"] * N / e2e["s"], "unit": "grid-point updates/s", "h2d_bytes_per_step": e2e["h2d"] / max(e2e["iters"], 1),
                 "d2h_bytes_per_step": e2e["d2h"] / max(e2e["iters"], 1), "seconds": e2e["s"],
                 "call": "PDHG_solver_oneiter(native callables, NumPy state in pinned host memory) -> NumPy: H2D phi/rho/alp, %d iterations, D2H "
                         "phi/rho/alp into pooled pinned memory (the transfers of the ~0.5 GB state bound this number at small step counts)" % e2e["iters"]}
  spin = SPINUP.get(a.workload, 0)
  cold_rf = roofline_of(pb, r, peak, peak_src, a.workload + "_cold")
  if spin:
    r2 = run_ours_block(pb, max(100, min(a.steps, 400)), warm, local, spinup=spin)     # (>= 100 iterations: the launch's state copy-in/out amortised)
    rf = roofline_of(pb, r2, peak, peak_src, a.workload)
    rf["window"] = "steady state: iterations %d..%d of the same solve (1 inner sweep per iteration)" % (r2["begin"], r2["begin"] + r2["iters"])
    rf["value"] = r2["iters"] * N / (r2["ms"] * 1e-3)
    rf["phase_us_per_iter"] = {k: round(v / max(r2["iters"], 1) * 1e3, 1) for k, v in r2["solver"].phase_times_ms().items() if k[0].isupper()}
    line["roofline"] = rf
    cold_rf["window"] = "the main line's cold window (fused dual passes move fewer bytes than the sweep-by-sweep algorithmic count: see dram_frac)"
    line["roofline_cold_window"] = cold_rf
  else:
    line["roofline"] = cold_rf
  line["cpu_baseline"] = cpu_baseline(pb)
  if not a.no_others:
    others = {}
    for nm, its in (("cfg3_tsp2", 2000), ("cfg2", 20000), ("cfg1", 3000), ("cfg5_tsp2", 100)):
      if nm != a.workload:
        try:
          others[nm] = secondary(nm, local, its)
        except Exception as ex:   # secondary lines never break the headline
          others[nm] = {"error": repr(ex)}
    for nm, nb in (("cfg1", None), ("cfg3_tsp2", 3)):
      try:
        others["time_to_tol_" + nm] = time_to_tol(nm, local, nb)
      except Exception as ex:
        others["time_to_tol_" + nm] = {"error": repr(ex)}
    try:
      others["cfg4_one_gpu"] = cfg4_sharded(0, 1, local, None)
    except Exception as ex:
      others["cfg4_one_gpu"] = {"error": repr(ex)}
    try:
      others["cfg4_cpu_baseline_all_cores"] = cpu_baseline_batched(nblk=1)
    except Exception as ex:
      others["cfg4_cpu_baseline_all_cores"] = {"error": repr(ex)}
    line["others"] = others
  print(json.dumps(line))
  return 0


def make_problem_cpu(name):
  """Problem description without touching the CUDA library (the reference arm must not load the product)."""
  from oracle import pdhg_numpy as orc
  egno, ndim, nx, ny, nt, tsp, epsl, stepsz = WORKLOADS[name]
  x_arr, bc, n_ctrl = orc.make_grid(egno, ndim, nx, ny, 2.0, 2.0)
  dt = 1.0 / (nt - 1)
  if ndim == 1:
    period, dspatial, nspatial = (2.0,), (2.0 / nx,), (nx,)
  else:
    period, dspatial, nspatial = (2.0, 2.0), (2.0 / nx, 2.0 / ny), (nx, ny)
  g = orc.set_up_J(egno, ndim, period)(x_arr)
  return dict(name=name, egno=egno, ndim=ndim, nx=nx, ny=ny, nt=nt, tsp=tsp, K=tsp - 1, epsl=epsl, stepsz=stepsz, n_ctrl=n_ctrl, bc=bc,
              x_arr=x_arr, dt=dt, dspatial=dspatial, nspatial=nspatial, g=g, N=(tsp - 1) * nx * ny)


if __name__ == "__main__":
  sys.exit(main())

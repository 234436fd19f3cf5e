"""bench.py — headline benchmark of the B200-native PDHG hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" = ONE outer PDHG iteration (primal phi update through the spectral preconditioner, extrapolation, dual
rho/alp sweeps, error reductions and exit tests; jaxsrc/utils/utils_pdhg_solver.py:51-88) on the named workload.
Headline workload (BASELINE.json configs[2], the config the north-star's HBM-roofline target is quoted on):
  cfg3_tsp65 = egno=1 ndim=2 epsl=0 nx=ny=256 nt=65 with --time_step_per_PDHG 65 (one space-time block,
  N = 64*256*256 = 4.19 M points, ~300 MB of state: larger than the 126 MB L2, i.e. HBM-bound).
Other workloads (secondary lines in the "others" key or via --workload): cfg3_tsp2 (reference default tsp=2,
L2-resident), cfg1 (README example, full solve), cfg2 (1-D viscous 640x161 at a stable step size).
N > 1: the path does not shard a single grid without a collective ("replicas only") — every rank runs
one independent instance of the workload (an epsl sweep), no data-path collective; value = sum over ranks.
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
  t1, it1 = oracle_iterations(pb, 1)
  n = int(max(1, min(200, budget_s / max(t1, 1e-6))))
  t, it = (t1, it1) if n == 1 else oracle_iterations(pb, n)
  return {"value": it * pb["N"] / t, "unit": "grid-point updates/s", "cores": 1, "kind": "port",
          "iters_per_s": it / t, "sample": "the first %d outer PDHG iteration(s) of block 0 of the same workload from the cold initial state, NumPy fp64 oracle, 1 process "
          "(JAX is not installable here, so the reference's own JAX-CPU path cannot be timed)" % it}


SPINUP = {"cfg3_tsp65": 600}    # untimed iterations before the timed region (steady-state inner-sweep count)


def batched_sample(device):
  """BASELINE configs[3] sample: independent 1-D instances (nx=1024, tsp=2, varied initial data / epsl) marched for the first
  4 of the 256 time blocks, one CTA per instance, 4 CTAs per SM's worth of instances."""
  import torch
  from pdhg_b200 import run_example as rx, set_fns
  B, nx, nt_full, nblk = 592, 1024, 257, 4
  rng = np.random.default_rng(0)
  A_, th, u = rng.uniform(0.5, 1.5, 4096)[:B], rng.uniform(0, 2 * np.pi, 4096)[:B], rng.uniform(0, 1, 4096)[:B]
  x_arr = rx.make_x_arr(1, nx, 1, 2.0, 2.0)
  g = A_[:, None] * np.sin(np.pi * x_arr[0, :, 0][None, :] + th[:, None])
  with contextlib.redirect_stdout(io.StringIO()):
    fns = set_fns.set_up_example_fns(1, 1, 0)
  T = nblk / (nt_full - 1)
  run = lambda: rx.solve_HJ_batch(1, 1, 1, 0.002 * u, fns, nx, 1, nblk + 1, 2.0, 2.0, T, x_arr, g, 70.0, 2, 0.1, 1000000, 10000, 1e-6, 0,
                                  device=device)
  run()
  t0 = time.perf_counter()
  phi, rho, alp, logs = run()
  t = time.perf_counter() - t0
  its = int(logs.iters.sum())
  return {"workload": "BASELINE configs[3] sample: %d of 4096 instances, nx=1024, first %d of 256 blocks, host buffers in/out" % (B, nblk),
          "instances": B, "total_iters": its, "seconds": t, "pdhg_iters_per_s": its / t, "grid_point_updates_per_s": its * nx / t,
          "all_converged": bool((logs.status == 0).all())}


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
    pb = make_problem_cpu(a.workload)
    t1, _ = oracle_iterations(pb, 1)
    for _ in range(min(warm, 3) - 1):
      oracle_iterations(pb, 1)
    n = int(max(1, min(a.steps, 150.0 / max(t1, 1e-6))))
    t, it = oracle_iterations(pb, n)
    val = it * pb["N"] / t
    cores = 1
    line = {"impl": "reference", "metric": "grid_point_updates_per_s", "value": val, "unit": "grid-point updates/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": t / it * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "pdhg_iters_per_s": it / t,
            "config": {"workload": DESCR[a.workload], "timed_steps": it},
            "cpu_baseline": {"value": val, "unit": "grid-point updates/s", "cores": cores, "kind": "port",
                             "sample": "%d of the requested %d outer PDHG iterations (block 0, same grid), NumPy fp64 oracle port of the "
                                       "reference, 1 process; the reference's JAX path cannot be installed in this image" % (it, a.steps)},
            "e2e": {"value": val, "unit": "grid-point updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0

  import torch
  import torch.distributed as dist
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device; the product path has no CPU fallback")
  barrier = None
  if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    barrier = lambda: (dist.barrier(), torch.cuda.synchronize())
  pb = make_problem(a.workload)
  sampler = ClockSampler(local) if rank == 0 else None
  # every rank: one independent instance (an epsl sweep: rank r uses epsl + 1e-3*r), no data-path collective
  # headline window: iterations [W, W+K) of the solve from its cold initial state — the same window the reference arm and the
  # cpu_baseline can afford on a CPU.  The steady state of the same solve (1 inner sweep / iteration) is reported separately.
  r = run_ours_block(pb, a.steps, warm, local, epsl_shift=1e-3 * rank, sampler=sampler, barrier=barrier, spinup=0)
  ms = r["ms"]
  iters = r["iters"]
  # end to end through the public API (host state in, host state out) on every rank; whole-job value = all iterations / max time
  if barrier:
    barrier()
  e2e = run_ours_e2e(pb, a.steps, local, r["state"])
  if world > 1:
    t = torch.tensor([ms, float(iters), r["kernel_ms"], e2e["s"], float(e2e["iters"])], dtype=torch.float64, device="cuda")
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    ms, total_iters = float(tmax[0]), float(tsum[1])
    e2e_s, e2e_iters = float(tmax[3]), float(tsum[4])
  else:
    total_iters = float(iters)
    e2e_s, e2e_iters = e2e["s"], float(e2e["iters"])
  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return 0
  N = pb["N"]
  value = total_iters * N / (ms * 1e-3)
  n_in = r["n_inner"] / max(iters, 1)
  by_launch = algorithmic_bytes_per_iter(pb["ndim"], pb["K"], pb["nx"], pb["ny"], n_in) * iters
  peak, peak_src = measured_peak_gbs()
  achieved = by_launch / (r["kernel_ms"] * 1e-3) / 1e9
  traffic = None
  tf = os.path.join(ROOT, "profiles", "traffic_%s.json" % a.workload)          # steady-state ncu capture
  tfc = os.path.join(ROOT, "profiles", "traffic_%s_cold.json" % a.workload)    # cold-window ncu capture
  if os.path.exists(tfc):
    try:
      traffic = json.load(open(tfc))["dram_bytes_per_iter"] * iters
    except Exception:
      traffic = None
  line = {"metric": "grid_point_updates_per_s", "value": value, "unit": "grid-point updates/s", "n_gpus": world, "steps": a.steps,
          "warmup": a.warmup, "ms_per_step": ms / max(iters, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f64", "data": "synthetic", "pdhg_iters_per_s": total_iters / (ms * 1e-3), "iters_timed": iters,
          "inner_sweeps_per_iter": n_in, "end_reason": r["end_reason"],
          "config": {"workload": DESCR[a.workload], "grid_points_per_iter": N, "stepsz_param": pb["stepsz"],
                     "timed_region": "iterations %d..%d of block 0 from the cold initial state (inner dual sweeps/iteration in this "
                                     "window: see inner_sweeps_per_iter; the solve settles at 1 after ~500 iterations: steady_state)"
                                     % (r["begin"], r["begin"] + iters),
                     "parallelism": "replicas x%d (independent instances, no collective)" % world,
                     "l2": "state ~%.0f MB > 126 MB L2 (inputs larger than L2, no flush needed)" % (N * 8 * 14 / 1e6)
                           if N * 8 * 14 > 126e6 else "state fits L2 (L2-resident regime; no flush: the iteration re-reads its own state)",
                     "kernel_path": "cooperative multi-CTA" if r["path"] == 2 else "single-CTA smem"},
          "gpu_launches": r["launches"], "clocks": r["clocks"],
          "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                       "kernel": "pdhg_coop_kernel" if r["path"] == 2 else "pdhg1d_cta_kernel", "kernel_ms": r["kernel_ms"],
                       "algorithmic_bytes_per_launch": by_launch, "peak_source": peak_src}}
  if e2e:
    line["e2e"] = {"value": e2e_iters * N / e2e_s, "unit": "grid-point updates/s", "h2d_bytes_per_step": e2e["h2d"] / max(e2e["iters"], 1),
                   "d2h_bytes_per_step": e2e["d2h"] / max(e2e["iters"], 1), "call": "PDHG_solver_oneiter(native callables, NumPy state in pinned host memory) -> NumPy: "
                   "H2D phi/rho/alp, %d iterations, D2H phi/rho/alp (per rank; value = all ranks' iterations / max time)" % e2e["iters"], "seconds": e2e_s}
  if world == 1:
    spin = SPINUP.get(a.workload, 0)
    if spin:
      k2 = max(20, min(a.steps, 200))
      r2 = run_ours_block(pb, k2, warm, local, spinup=spin)
      n2 = r2["n_inner"] / max(r2["iters"], 1)
      by2 = algorithmic_bytes_per_iter(pb["ndim"], pb["K"], pb["nx"], pb["ny"], n2) * r2["iters"]
      ach2 = by2 / (r2["kernel_ms"] * 1e-3) / 1e9
      tr2 = None
      if os.path.exists(tf):
        try:
          tr2 = json.load(open(tf))["dram_bytes_per_iter"] * r2["iters"]
        except Exception:
          tr2 = None
      line["steady_state"] = {"timed_region": "iterations %d..%d of the same solve" % (r2["begin"], r2["begin"] + r2["iters"]),
                              "value": r2["iters"] * N / (r2["ms"] * 1e-3), "unit": "grid-point updates/s", "ms_per_step": r2["ms"] / max(r2["iters"], 1),
                              "pdhg_iters_per_s": r2["iters"] / (r2["ms"] * 1e-3), "inner_sweeps_per_iter": n2,
                              "roofline": {"bound": "hbm", "achieved": ach2, "peak": peak, "unit": "GB/s", "frac": ach2 / peak, "traffic": tr2,
                                           "kernel_ms": r2["kernel_ms"], "algorithmic_bytes_per_launch": by2}}
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
        others["cfg4_sample"] = batched_sample(local)
      except Exception as ex:
        others["cfg4_sample"] = {"error": repr(ex)}
      line["others"] = others
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()
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

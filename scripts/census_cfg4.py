"""Stopping-iteration census on BASELINE configs[3]-style instances: B instances x nblk time blocks, GPU
(both 1-D kernel families) against the NumPy oracle.  The reference's stopping rule compares relative CHANGES with eps = 1e-6;
where those hover around eps the stopping iteration depends on rounding, so it is reported per instance-block: how many agree
exactly, and how far the solutions are where they do not.

    python scripts/census_cfg4.py [B=64] [nblk=4] [out.json]
"""
import contextlib, io, json, multiprocessing as mp, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pdhg-optimal-control_b200"))
import numpy as np
import bench


def oracle_instance(args):
  b, nblk, nx, nt_full = args
  from oracle import pdhg_numpy as orc
  g, epsl = bench.cfg4_inputs(b + 1, nx)
  x_arr, bc, n_ctrl = orc.make_grid(1, 1, nx, 1, 2.0, 2.0)
  info = {}
  res, _ = orc.solve_HJ(1, n_ctrl, 1, float(epsl[b]), orc.set_up_example_fns(1, 1, 0), nx, 1, nblk + 1, 2.0, 2.0, nblk / (nt_full - 1.0), x_arr, 70.0, 2,
                        0.1, 1000000, 10 ** 9, 1e-6, bc, g=g[b:b + 1], info=info)
  return info["block_iters"], np.asarray(res[0][1]), np.asarray(res[0][2])


def main():
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
  nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 4
  out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "census_cfg4.json")
  nx, nt_full = 1024, 257
  from pdhg_b200 import run_example as rx, set_fns
  g, epsl = bench.cfg4_inputs(B, nx)
  x_arr = rx.make_x_arr(1, nx, 1, 2.0, 2.0)
  with contextlib.redirect_stdout(io.StringIO()):
    fns = set_fns.set_up_example_fns(1, 1, 0)
  T = nblk / (nt_full - 1.0)
  gpu = {}
  for name, env in (("k1_register_resident", {}), ("cta_shared_memory", {"PDHG_NO_K1": "1"})):
    os.environ.pop("PDHG_NO_K1", None)
    os.environ.update(env)
    phi, rho, alp, logs = rx.solve_HJ_batch(1, 1, 1, epsl, fns, nx, 1, nblk + 1, 2.0, 2.0, T, x_arr, g, 70.0, 2, 0.1, 1000000, 10 ** 9, 1e-6, 0)
    gpu[name] = (logs.iters.copy(), phi.copy(), rho.copy())
  os.environ.pop("PDHG_NO_K1", None)
  t0 = time.time()
  with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, B)) as pool:
    orc_res = pool.map(oracle_instance, [(b, nblk, nx, nt_full) for b in range(B)])
  t_orc = time.time() - t0
  rel = lambda a, b: float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
  report = {"instances": B, "blocks": nblk, "oracle_seconds": t_orc, "kernels": {}}
  for name, (it, phi, rho) in gpu.items():
    same, diffs, worst_same, worst_diff = 0, [], 0.0, 0.0
    for b in range(B):
      ito, phio, rhoo = orc_res[b]
      for k in range(nblk):
        e = max(rel(phi[b, k + 1], phio[k + 1]), rel(rho[b, k], rhoo[k]))
        if int(it[b, k]) == int(ito[k]):
          same += 1
          # (a later block inherits the difference of an earlier block whose stopping iteration differed)
          worst_same = max(worst_same, e)
        else:
          diffs.append({"instance": b, "block": k, "gpu_iters": int(it[b, k]), "oracle_iters": int(ito[k]), "rel_linf_phi_rho": e})
          worst_diff = max(worst_diff, e)
    report["kernels"][name] = {"instance_blocks": B * nblk, "identical_stopping_iteration": same, "fraction_identical": same / (B * nblk),
                               "worst_rel_linf_where_identical": worst_same, "worst_rel_linf_where_different": worst_diff, "different": diffs}
  i1, i2 = gpu["k1_register_resident"][0], gpu["cta_shared_memory"][0]
  report["k1_vs_cta_identical_fraction"] = float(np.mean(i1 == i2))
  os.makedirs(os.path.dirname(out), exist_ok=True)
  json.dump(report, open(out, "w"), indent=1)
  for name, r in report["kernels"].items():
    print(name, "identical %d / %d (%.3f), worst rel-Linf identical %.2e, different %.2e, n_different %d" % (
      r["identical_stopping_iteration"], r["instance_blocks"], r["fraction_identical"], r["worst_rel_linf_where_identical"],
      r["worst_rel_linf_where_different"], len(r["different"])))
  print("k1 vs cta identical fraction %.3f; oracle %.0f s" % (report["k1_vs_cta_identical_fraction"], t_orc))


if __name__ == "__main__":
  main()

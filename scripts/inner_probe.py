"""Scratch: how the inner dual sweep count evolves over a block solve (cfg3_tsp65)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pdhg-optimal-control_b200"))
import numpy as np, torch, bench
from pdhg_b200.update_fns_in_pdhg import get_solver
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3_tsp65"
stepsz = float(sys.argv[2]) if len(sys.argv) > 2 else None
pb = bench.make_problem(name)
if stepsz: pb["stepsz"] = stepsz
s = get_solver(pb["fns"], pb["nspatial"], pb["K"], pb["bc"], pb["dt"], pb["dspatial"], 70.0, pb["x_arr"], nblocks=1, max_rec=4)
K, nsp, A = pb["K"], tuple(pb["nspatial"]), 2 * pb["ndim"]
dev = torch.device("cuda", 0)
g = torch.from_numpy(np.ascontiguousarray(pb["g"])).to(dev)
phi = g.expand((K + 1,) + nsp).contiguous(); rho = torch.full((K,) + nsp, 70.0, dtype=torch.float64, device=dev)
alp = torch.zeros((A, K) + nsp + (pb["n_ctrl"],), dtype=torch.float64, device=dev)
marks = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "50,200,500,1000,2000,4000,8000".split(","))]
i0 = 0
for m in marks:
  po, ro, ao = torch.empty_like(phi), torch.empty_like(rho), torch.empty_like(alp)
  logs = s.solve_block_dev(phi.data_ptr(), rho.data_ptr(), alp.data_ptr(), pb["epsl"], pb["stepsz"], 10**6, i0, m, 0, po.data_ptr(), ro.data_ptr(), ao.data_ptr(), None)
  it = int(logs.iters[0, 0]); n = it - i0
  print("iters %d..%d: inner/iter %.2f  us/iter %.1f  end_reason %d err %s" % (i0, it, logs.inner_total[0] / max(n, 1), s.last_kernel_ms * 1e3 / max(n, 1), logs.end_reason[0, 0], logs.errlog[0, 0, 0, :2]), flush=True)
  phi, rho, alp = po, ro, ao
  if logs.end_reason[0, 0] != 3: break
  i0 = it

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for args in "cfg3_tsp2 2000 0" "cfg3_tsp2 2000 2000" "cfg5_tsp2 100 0"; do
  PDHG_PROFILE=1 timeout 200 python scripts/phase_breakdown.py $args 2>&1 | tail -1 | cut -c1-900
  PDHG_PROFILE= timeout 200 python - <<PY 2>&1 | tail -1
import os, sys
os.environ.pop("PDHG_PROFILE", None)
sys.path.insert(0, ".")
import bench
a = "$args".split()
pb = bench.make_problem(a[0])
r = bench.run_ours_block(pb, int(a[1]), 3, 0, spinup=int(a[2]))
print("no-profile us/iter", r["ms"] / r["iters"] * 1e3, "sweeps/iter", r["n_inner"] / r["iters"])
PY
done

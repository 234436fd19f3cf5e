#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r3k_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print({k: d[k] for k in ("value", "ms_per_step", "steps", "warmup", "gpu_launches")}, d["e2e"]["value"])
        rf = d["roofline"]; print({k: rf.get(k) for k in ("achieved", "frac", "traffic", "dram_frac", "ms_per_iter", "phase_us_per_iter")})
        print(list(d.get("others", {}).keys()))
PY
tail -2 gpurun_out/r3k_bench.err | cut -c1-300

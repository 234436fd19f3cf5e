#!/bin/bash
# N-rank bench line (sharded configs[3] + slab line in others), as the driver launches it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3c_bench_n$N.json 2> gpurun_out/r3c_bench_n$N.err
echo "rc=$?"
python - <<PY
import json
for l in open("gpurun_out/r3c_bench_n$N.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "pdhg_iters_per_s")}, d["e2e"]["value"])
        o = d.get("others", {})
        print(json.dumps(o.get("cfg5_slab"))[:900])
        print(json.dumps(o.get("cfg4_one_gpu"))[:300])
PY
tail -3 gpurun_out/r3c_bench_n$N.err | cut -c1-300

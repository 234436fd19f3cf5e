#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2r_bench_n$N.json 2> gpurun_out/r2r_bench_n$N.err
echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/slab_prof.py 2048 20 > gpurun_out/r2r_slab_prof_n$N.txt 2>&1
grep "^{" gpurun_out/r2r_slab_prof_n$N.txt | cut -c1-400
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/r2r_bench_n$N.json") if l.startswith("{")][-1]
print("N", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "kernel_s", d["ms_per_step"]/1e3)
print("slab", d["others"].get("cfg5_slab"))
print("one_gpu", d["others"].get("cfg4_one_gpu",{}).get("value_device"))
PY

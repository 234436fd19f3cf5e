#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
bash scripts/gpu/r3c.sh $N
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 scripts/slab_host_cost.py 2048 > gpurun_out/r3d_host_cost_n$N.txt 2>&1
grep "^{" gpurun_out/r3d_host_cost_n$N.txt; tail -5 gpurun_out/r3d_host_cost_n$N.txt | grep -v "^{" | cut -c1-300

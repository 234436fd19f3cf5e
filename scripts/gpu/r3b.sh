#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-4}
export PDHG_SLAB_GROUP=symm
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/slab_bench.py 512 20 > gpurun_out/r3b_parity_n$N.txt 2>&1
grep "^{" gpurun_out/r3b_parity_n$N.txt | cut -c1-700; tail -4 gpurun_out/r3b_parity_n$N.txt | grep -v "^{" | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/slab_prof.py 2048 40 > gpurun_out/r3b_prof_n$N.txt 2>&1
grep "^{" gpurun_out/r3b_prof_n$N.txt | cut -c1-520; tail -4 gpurun_out/r3b_prof_n$N.txt | grep -v "^{" | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 scripts/slab_host_cost.py 2048 > gpurun_out/r3b_host_cost_n$N.txt 2>&1
grep "^{" gpurun_out/r3b_host_cost_n$N.txt; tail -5 gpurun_out/r3b_host_cost_n$N.txt | grep -v "^{" | cut -c1-300

#!/bin/bash
# slab decomposition: symmetric-memory exchanges against NCCL exchanges (parity at 512^2, sections + rate at 2048^2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
for kind in symm nccl; do
  export PDHG_SLAB_GROUP=$kind
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/slab_bench.py 512 20 > gpurun_out/r2v_parity_${kind}_n$N.txt 2>&1
  grep "^{" gpurun_out/r2v_parity_${kind}_n$N.txt | cut -c1-700; tail -4 gpurun_out/r2v_parity_${kind}_n$N.txt | grep -v "^{" | cut -c1-300
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/slab_prof.py 2048 20 > gpurun_out/r2v_prof_${kind}_n$N.txt 2>&1
  grep "^{" gpurun_out/r2v_prof_${kind}_n$N.txt | cut -c1-520; tail -4 gpurun_out/r2v_prof_${kind}_n$N.txt | grep -v "^{" | cut -c1-300
done

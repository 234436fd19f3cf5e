#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
export PDHG_SLAB_GROUP=symm PDHG_SLAB_PROF_QUICK=1
for fx in B bwd 0; do
PDHG_SLAB_FUSED_XCH=$fx timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/slab_prof.py 2048 60 > gpurun_out/r3i_prof_${fx}_n$N.txt 2>&1
echo "mode $fx"; grep "^{" gpurun_out/r3i_prof_${fx}_n$N.txt | cut -c1-300; tail -3 gpurun_out/r3i_prof_${fx}_n$N.txt | grep -v "^{" | cut -c1-300
done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
export PDHG_SLAB_GROUP=symm PDHG_SLAB_TRACE=1
for ahead in 0 1; do
PDHG_SLAB_AHEAD=$ahead timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/slab_bench.py 512 12 > gpurun_out/r3a_trace_ahead$ahead.txt 2>&1
grep "^slab it\|^{" gpurun_out/r3a_trace_ahead$ahead.txt | cut -c1-330
done

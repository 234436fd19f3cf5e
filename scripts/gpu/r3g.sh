#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -u -m pytest tests/test_gpu_slab.py -m gpu -x -q --timeout 300 > gpurun_out/r3g_tests.txt 2>&1
tail -5 gpurun_out/r3g_tests.txt | cut -c1-250
export PDHG_SLAB_GROUP=symm
for fx in 1 0; do
export PDHG_SLAB_FUSED_XCH=$fx
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/slab_bench.py 512 20 > gpurun_out/r3g_parity_fx${fx}_n$N.txt 2>&1
grep "^{" gpurun_out/r3g_parity_fx${fx}_n$N.txt | cut -c1-700; tail -4 gpurun_out/r3g_parity_fx${fx}_n$N.txt | grep -v "^{" | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/slab_prof.py 2048 40 > gpurun_out/r3g_prof_fx${fx}_n$N.txt 2>&1
grep "^{" gpurun_out/r3g_prof_fx${fx}_n$N.txt | cut -c1-520; tail -4 gpurun_out/r3g_prof_fx${fx}_n$N.txt | grep -v "^{" | cut -c1-300
done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
export PDHG_SLAB_GROUP=symm
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/slab_bench.py 512 20 > gpurun_out/r3e_parity_n$N.txt 2>&1
grep "^{" gpurun_out/r3e_parity_n$N.txt | cut -c1-700; tail -4 gpurun_out/r3e_parity_n$N.txt | grep -v "^{" | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/slab_prof.py 2048 40 > gpurun_out/r3e_prof_n$N.txt 2>&1
grep "^{" gpurun_out/r3e_prof_n$N.txt | cut -c1-520; tail -4 gpurun_out/r3e_prof_n$N.txt | grep -v "^{" | cut -c1-300
bash scripts/gpu/r3c.sh $N
cp gpurun_out/r3c_bench_n$N.json gpurun_out/r3e_bench_n$N.json

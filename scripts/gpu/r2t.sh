#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/slab_prof.py 2048 20 > gpurun_out/r2t_slab_prof_n$N.txt 2>&1
grep "^{" gpurun_out/r2t_slab_prof_n$N.txt | cut -c1-420
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/slab_bench.py 512 20 > gpurun_out/r2t_slab_parity_n$N.txt 2>&1
grep "^{" gpurun_out/r2t_slab_parity_n$N.txt | cut -c1-900; tail -3 gpurun_out/r2t_slab_parity_n$N.txt | grep -v "^{" | cut -c1-300

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2m_bench_n$N.err
tail -c 3000 gpurun_out/r2m_bench_n$N.json

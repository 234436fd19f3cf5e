#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cd scripts/micro
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dual_tma_bench dual_tma_bench.cu > /dev/null 2>&1 && timeout 200 /tmp/dual_tma_bench > ../../gpurun_out/r2c_dual_tma.txt 2>&1

cat ../../gpurun_out/r2c_dual_tma.txt

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python scripts/phase_ncu.py > gpurun_out/r2e_phase_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pdhg_coop -s 9 -c 1 -o gpurun_out/r2e_D -f python scripts/phase_ncu.py > gpurun_out/r2e_ncu_D.log 2>&1
cd scripts/micro
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/dual_tma_bench dual_tma_bench.cu > /dev/null 2>&1
/tmp/dual_tma_bench > ../../gpurun_out/r2e_micro_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_direct -s 2 -c 1 -o ../../gpurun_out/r2e_kdirect -f /tmp/dual_tma_bench > ../../gpurun_out/r2e_ncu_k.log 2>&1
ls -la ../../gpurun_out/

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 scripts/symm_probe.py > gpurun_out/r2u_symm.txt 2>&1
grep -E "rank|Error|error" gpurun_out/r2u_symm.txt | tail -12 | cut -c1-400

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_slab.py -m gpu -x -q -v -s --timeout 300 > gpurun_out/r2w_tests.txt 2>&1
tail -25 gpurun_out/r2w_tests.txt | cut -c1-250

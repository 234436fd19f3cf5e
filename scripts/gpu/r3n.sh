#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 80 python -u -m pytest tests/test_gpu_variants.py -m gpu -q --timeout 70 -k "cfg3_blocks012 or cfg2_stepsz01" > gpurun_out/r3n_tests.txt 2>&1
tail -3 gpurun_out/r3n_tests.txt | cut -c1-250

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 100 python -u -m pytest tests/test_gpu_slab.py -m gpu -q --timeout 90 > gpurun_out/r3m_tests.txt 2>&1
tail -3 gpurun_out/r3m_tests.txt | cut -c1-250

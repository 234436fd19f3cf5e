#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_variants.py -v -m gpu -x --timeout=150 -k "single_pass or fused or tma" > gpurun_out/r2g_tests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/r2g_tests.txt
grep -E "PASS|FAIL|ERROR|Timeout|passed|failed|rc=" gpurun_out/r2g_tests.txt | tail -40

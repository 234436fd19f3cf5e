#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -u -m pytest tests/test_traj.py -v -m gpu --timeout=120 > gpurun_out/r2n_tests.txt 2>&1
echo "tests rc=$?"
grep -E "PASS|FAIL|ERROR|passed|failed|Error|assert" gpurun_out/r2n_tests.txt | tail -25

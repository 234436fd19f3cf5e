#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -u -m pytest tests -q -m gpu -x --timeout=300 > gpurun_out/r2l_tests.txt 2>&1
echo "tests rc=$?"
tail -5 gpurun_out/r2l_tests.txt

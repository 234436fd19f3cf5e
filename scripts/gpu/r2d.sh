#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
PDHG_NO_LEAN=1 timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2d_probe_nolean.txt 2>&1
timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2d_probe_lean.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_variants.py -q -m gpu -x -k "not headline" > gpurun_out/r2d_tests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests.txt
tail -3 gpurun_out/r2d_tests.txt
cat gpurun_out/r2d_probe_*.txt

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
PDHG_NO_BSLAB=1 timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2f_probe_nobslab.txt 2>&1
timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2f_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_variants.py -q -m gpu -x -k "single_pass or headline or fused" > gpurun_out/r2f_tests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/r2f_tests.txt
tail -15 gpurun_out/r2f_tests.txt
cat gpurun_out/r2f_probe*.txt

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -u -m pytest tests/test_gpu_variants.py -v -m gpu -x --timeout=150 -k "single_pass" > gpurun_out/r2h_tests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/r2h_tests.txt
grep -E "PASS|FAIL|ERROR|Timeout|passed|failed|rc=" gpurun_out/r2h_tests.txt | tail -12
timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2h_probe.txt 2>&1
cat gpurun_out/r2h_probe.txt

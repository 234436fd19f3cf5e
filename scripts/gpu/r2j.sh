#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -u -m pytest tests/test_gpu_variants.py -v -m gpu -x --timeout=150 -k "hybrid or single_pass" > gpurun_out/r2j_tests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/r2j_tests.txt
grep -E "PASS|FAIL|ERROR|Timeout|passed|failed|rc=|assert" gpurun_out/r2j_tests.txt | tail -14
timeout 200 python scripts/phase_probe.py 100 > gpurun_out/r2j_probe.txt 2>&1
PDHG_NO_HYB=1 timeout 200 python scripts/phase_probe.py 100 > gpurun_out/r2j_probe_nohyb.txt 2>&1
PDHG_TMA=1 timeout 200 python scripts/phase_probe.py 100 > gpurun_out/r2j_probe_tma.txt 2>&1
head -1 gpurun_out/r2j_probe.txt gpurun_out/r2j_probe_nohyb.txt gpurun_out/r2j_probe_tma.txt

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2k_bench.err
timeout 900 python -u -m pytest tests -q -m gpu -x --timeout=300 > gpurun_out/r2k_tests.txt 2>&1
echo "tests rc=$?"
tail -5 gpurun_out/r2k_tests.txt

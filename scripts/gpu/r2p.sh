#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_gpu_slab.py tests/test_gpu_variants.py -q -m gpu -x --timeout=200 -k "slab or fused or long_transform" > gpurun_out/r2p_tests.txt 2>&1
echo "tests rc=$?"
tail -4 gpurun_out/r2p_tests.txt
python - <<'PY' > gpurun_out/r2p_cfg5.txt 2>&1
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "pdhg-optimal-control_b200"))
import bench
print(bench.secondary("cfg5_tsp2", 0, 100))
PY
cat gpurun_out/r2p_cfg5.txt | tail -2

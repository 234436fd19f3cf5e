#!/bin/bash
# round 2, GPU call A: phase timers old vs TMA dual sweep, dual-sweep launch-shape micro-benchmark, variant parity tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
for tma in 0 1; do
  PDHG_TMA=$tma timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2a_probe_tma$tma.txt 2>&1
done
( cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dual_bench dual_bench.cu > /dev/null 2>&1 && timeout 120 /tmp/dual_bench ) > gpurun_out/r2a_dual_bench.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_variants.py -q -m gpu -x --durations=15 > gpurun_out/r2a_variants.txt 2>&1
echo "variants rc=$?" >> gpurun_out/r2a_variants.txt
tail -5 gpurun_out/r2a_variants.txt
cat gpurun_out/r2a_probe_tma0.txt gpurun_out/r2a_probe_tma1.txt

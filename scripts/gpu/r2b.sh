#!/bin/bash
# round 2, GPU call B: asm vs plain load/store builds (phase timers), TMA ring micro-benchmark, ncu of the single phases
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=pdhg-optimal-control_b200/lib
for v in asm b200; do
  PDHG_B200_LIB=$PWD/$L/libpdhg_$v.so PDHG_TMA=0 timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2b_probe_$v.txt 2>&1
done
PDHG_TMA=1 timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2b_probe_b200_tma1.txt 2>&1
( cd scripts/micro && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/stream_bench stream_bench.cu > /dev/null 2>&1 && timeout 120 /tmp/stream_bench ) > gpurun_out/r2b_stream_bench.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py -q -m gpu -x -k "not headline and not cfg1_readme" > gpurun_out/r2b_tests.txt 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.txt
PDHG_TMA=0 python scripts/phase_ncu.py > gpurun_out/r2b_phase_plain.log 2>&1 && \
PDHG_TMA=0 ncu --set full --clock-control none --import-source on -k regex:pdhg_coop -s 4 -c 8 -o gpurun_out/r2b_phases -f python scripts/phase_ncu.py > gpurun_out/r2b_phase_ncu.log 2>&1
tail -3 gpurun_out/r2b_tests.txt
cat gpurun_out/r2b_probe_*.txt

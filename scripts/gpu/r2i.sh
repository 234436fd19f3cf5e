#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=$PWD/pdhg-optimal-control_b200/lib
timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2i_probe_A.txt 2>&1
PDHG_B200_LIB=$L/libpdhg_tb.so timeout 300 python scripts/phase_probe.py 100 > gpurun_out/r2i_probe_B.txt 2>&1
head -1 gpurun_out/r2i_probe_A.txt; head -1 gpurun_out/r2i_probe_B.txt

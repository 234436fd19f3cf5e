#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -u -m pytest tests/test_gpu_slab.py tests/test_gpu_integration_doc.py -m gpu -q --timeout 300 > gpurun_out/r3l_tests.txt 2>&1
tail -4 gpurun_out/r3l_tests.txt | cut -c1-250
timeout 300 python scripts/traffic_capture.py 100 > gpurun_out/r3l_traffic_plain.log 2>&1 && \
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pdhg_coop --csv --log-file gpurun_out/r02_traffic.csv python scripts/traffic_capture.py 100 > gpurun_out/r3l_traffic_ncu.log 2>&1
tail -2 gpurun_out/r3l_traffic_plain.log

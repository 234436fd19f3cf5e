#!/bin/bash
# final pass of round 2: full GPU test suite, DRAM traffic of the bench windows, launch list of the bench command, per-phase ncu capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -u -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r3j_tests.txt 2>&1
tail -6 gpurun_out/r3j_tests.txt | cut -c1-250
timeout 300 python scripts/traffic_capture.py 100 > gpurun_out/r3j_traffic_plain.log 2>&1 && \
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pdhg_coop --csv --log-file gpurun_out/r02_traffic.csv python scripts/traffic_capture.py 100 > gpurun_out/r3j_traffic_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-others > gpurun_out/r3j_bench_plain.json 2> gpurun_out/r3j_bench_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 5 --no-others > gpurun_out/r3j_bench_ncu.log 2>&1
timeout 300 python scripts/phase_ncu.py > gpurun_out/r3j_phase_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:pdhg_coop -s 4 -c 8 --csv --page raw --log-file gpurun_out/r02_phases_raw.csv python scripts/phase_ncu.py > gpurun_out/r3j_phase_ncu.log 2>&1
cat gpurun_out/r3j_traffic_plain.log | tail -3
cut -c1-600 gpurun_out/r3j_bench_plain.json

#!/bin/bash
# final profiling pass: DRAM traffic of the steady / cold windows, launch list of the bench command, per-phase ncu --set full
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python scripts/traffic_capture.py 100 > gpurun_out/r2s_traffic_plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pdhg_coop --csv --log-file gpurun_out/r02_traffic.csv python scripts/traffic_capture.py 100 > gpurun_out/r2s_traffic_ncu.log 2>&1
python bench.py --steps 20 --warmup 5 --no-others > gpurun_out/r2s_bench_plain.json 2> gpurun_out/r2s_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 5 --no-others > gpurun_out/r2s_bench_ncu.log 2>&1
python scripts/phase_ncu.py > gpurun_out/r2s_phase_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:pdhg_coop -s 4 -c 8 --csv --page raw --log-file gpurun_out/r02_phases_raw.csv python scripts/phase_ncu.py > gpurun_out/r2s_phase_ncu.log 2>&1
ls -la gpurun_out | tail -12
cat gpurun_out/r2s_traffic_plain.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python scripts/census_cfg4.py 64 4 gpurun_out/r02_census_cfg4.json > gpurun_out/r2o_census.txt 2>&1
cat gpurun_out/r2o_census.txt | tail -5

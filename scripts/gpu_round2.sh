#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "multigrid or tile or sdf_2d" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
tail -30 gpurun_out/pytest_gpu2.log
FI_B200_TRACE=1 timeout 300 python scripts/mg_explore.py 512 1000000 3:12 > gpurun_out/mg_trace.log 2>&1; tail -60 gpurun_out/mg_trace.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mg_launches.csv python scripts/profile_mg.py 512 2 > gpurun_out/ncu_mg.log 2>&1; tail -3 gpurun_out/ncu_mg.log

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu3.log
tail -45 gpurun_out/pytest_gpu3.log
FI_B200_TRACE=1 timeout 300 python scripts/mg_explore.py 256,512 1000000 3:12,2:12,4:12,3:30 > gpurun_out/mg_trace3.log 2>&1; grep -v "^\[fi" gpurun_out/mg_trace3.log; grep build_multigrid gpurun_out/mg_trace3.log | head -3
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mg_launches3.csv python scripts/profile_mg.py 512 2 > gpurun_out/ncu_mg3.log 2>&1; tail -3 gpurun_out/ncu_mg3.log

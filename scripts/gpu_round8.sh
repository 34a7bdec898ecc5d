#!/bin/bash
# 1-GPU visit: GPU suite (minus the three CPU-bound exact-solve cases, re-run in the final visit), data-term kernel A/B,
# default bench line, multigrid time-to-1e-6 with the device-side coarsest inverse (coarsest <= 600 cells vs 150).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --durations=8 -k "not (sizes4 or orders2)" > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu8.log
tail -15 gpurun_out/pytest_gpu8.log
timeout 200 python scripts/time_iters.py 256,512 > gpurun_out/time_iters8_split.jsonl 2>&1; grep '"fast": true' gpurun_out/time_iters8_split.jsonl
FI_B200_DATA_TERM=cell timeout 200 python scripts/time_iters.py 512 > gpurun_out/time_iters8_cell.jsonl 2>&1; grep '"fast": true' gpurun_out/time_iters8_cell.jsonl
timeout 200 python scripts/mg_explore.py 256,512 1000000 3:12,2:12 > gpurun_out/mg_explore8_c600.jsonl 2>&1; grep -v "^\[fi" gpurun_out/mg_explore8_c600.jsonl | tail -8
FI_B200_MG_COARSEST=4100 timeout 200 python scripts/mg_explore.py 512 1000000 3:12 > gpurun_out/mg_explore8_c4100.jsonl 2>&1; grep -v "^\[fi" gpurun_out/mg_explore8_c4100.jsonl | tail -4
timeout 400 python bench.py > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -c 1500 gpurun_out/bench8.json; tail -5 gpurun_out/bench8.err

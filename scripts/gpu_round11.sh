#!/bin/bash
# First GPU visit of the next round (1 GPU, ~12 min of box time): everything added after round 1's GPU budget ran out.
#   1. the whole GPU suite (the iso-surface tests have never run on a GPU; slab multigrid ran on 1 rank only)
#   2. default bench line (N=1) + reference arm
#   3. ncu launch list of the bench step and one full capture of the dominant kernels -> profiles/ summaries
# Usage: gpurun --timeout 900 -- 'bash scripts/gpu_round11.sh'
set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu11.log
tail -15 gpurun_out/pytest_gpu11.log
timeout 300 python bench.py > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -c 1500 gpurun_out/bench11.json; tail -5 gpurun_out/bench11.err
timeout 200 python bench.py --impl reference > gpurun_out/bench11_reference.json 2> gpurun_out/bench11_reference.err; tail -c 600 gpurun_out/bench11_reference.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches11.csv \
    python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu-baseline --no-time-to-tol > gpurun_out/bench11_under_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/launches11.csv > gpurun_out/launches11.md 2>&1; head -30 gpurun_out/launches11.md

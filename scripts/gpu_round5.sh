#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu5.log
tail -40 gpurun_out/pytest_gpu5.log
timeout 300 python bench.py --no-cpu-baseline --no-time-to-tol > gpurun_out/bench5.json 2> gpurun_out/bench5.err; tail -c 2500 gpurun_out/bench5.json; tail -5 gpurun_out/bench5.err

#!/bin/bash
# r1c visit 1: validate the table-driven multigrid + fused Chebyshev commit, full GPU suite, default bench line.
set -x
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu4.log
tail -30 gpurun_out/pytest_gpu4.log
timeout 500 python bench.py > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -c 4000 gpurun_out/bench4.json; tail -5 gpurun_out/bench4.err
timeout 200 python scripts/mg_explore.py 256,512 1000000 3:12 > gpurun_out/mg_explore4.jsonl 2>&1; grep -v "^\[fi" gpurun_out/mg_explore4.jsonl | tail -8

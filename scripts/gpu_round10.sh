#!/bin/bash
# Last minutes of the round-1 GPU budget: the unvalidated tests first (slab multigrid on a 1-rank communicator, the
# refactored single-GPU multigrid), then whatever else fits.
set -x
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_dist.py tests/test_gpu_solve.py -m gpu -q -x -k "slab or multigrid" > gpurun_out/pytest10.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest10.log
tail -12 gpurun_out/pytest10.log
timeout 120 python bench.py --steps 100 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -c 1200 gpurun_out/bench10.json; tail -3 gpurun_out/bench10.err

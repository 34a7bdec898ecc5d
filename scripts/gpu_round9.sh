#!/bin/bash
# 1-GPU visit: the slab-sharded multigrid path on a 1-rank communicator (new), then the whole GPU suite (mg.cu was
# refactored), 1-rank slab-MG timing at bench sizes, default bench line.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dist.py -m gpu -q -x > gpurun_out/pytest_dist9.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dist9.log
tail -25 gpurun_out/pytest_dist9.log
timeout 300 python scripts/slab_mg_one_rank.py 256,512 > gpurun_out/slab_mg_one_rank9.jsonl 2>&1; grep -v "^\[fi" gpurun_out/slab_mg_one_rank9.jsonl | tail -8
timeout 200 python scripts/slab_mg_one_rank.py 256 300000 > gpurun_out/slab_mg_one_rank9_g300k.jsonl 2>&1; grep -v "^\[fi" gpurun_out/slab_mg_one_rank9_g300k.jsonl | tail -4
timeout 600 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu9.log
tail -15 gpurun_out/pytest_gpu9.log
timeout 400 python bench.py > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -c 1500 gpurun_out/bench9.json; tail -5 gpurun_out/bench9.err

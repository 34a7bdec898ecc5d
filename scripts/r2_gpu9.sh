#!/bin/bash
# Round 2, last single-GPU visit: what the driver runs at round end (whole GPU suite, smoke(), the default bench line), then the
# time-to-1e-6 table and a phase trace with the cheaper hierarchy setup (dense coarsest matrix written directly, one-launch
# Gauss-Jordan steps, power iterations without per-round synchronisation).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2q_pytest.log
tail -5 gpurun_out/r2q_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; tail -2 gpurun_out/r2q_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; tail -c 400 gpurun_out/r2q_bench.json; tail -3 gpurun_out/r2q_bench.err
timeout 200 python scripts/r2_time_to_tol.py 512 C4 C3 C2 > gpurun_out/r2q_time_to_tol.jsonl 2> gpurun_out/r2q_time_to_tol.err
cut -c 1-300 gpurun_out/r2q_time_to_tol.jsonl | grep '"rep": 1'; tail -3 gpurun_out/r2q_time_to_tol.err
FI_B200_TRACE=1 timeout 120 python scripts/r2_time_to_tol.py 512 C2 > gpurun_out/r2q_trace_time_to_tol.txt 2>&1; grep -v "^{" gpurun_out/r2q_trace_time_to_tol.txt | tail -44 | cut -c 1-160

#!/bin/bash
# One GPU-box visit: parity tests, both bench arms, ncu launch list + full capture, kernel timings.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
timeout 400 python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 300 python scripts/time_iters.py 256,512 > gpurun_out/time_iters.jsonl 2>&1
timeout 300 python scripts/mg_explore.py 256,512 1000000 3:12 > gpurun_out/mg_explore.jsonl 2>&1; cat gpurun_out/mg_explore.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu-baseline --no-time-to-tol > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stencil3d_tma|pcg_update|apply_blocks" -s 12 -c 6 -f -o gpurun_out/full python scripts/profile_step.py 512 12 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out

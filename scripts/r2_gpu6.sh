#!/bin/bash
# Round 2, GPU visit 6 (1 GPU): 7-row vs 8-row stencil tiles; warm phase trace of a bench step; suite; bench.
set -x
mkdir -p gpurun_out
for ty in 8 7 0; do
  FI_B200_STENCIL_TY=$ty timeout 200 python scripts/time_iters.py 512,256 1000000 200 > gpurun_out/r2g_time_iters_ty$ty.jsonl 2> gpurun_out/r2g_time_iters_ty$ty.err
  grep '"fast": true' gpurun_out/r2g_time_iters_ty$ty.jsonl | cut -c 1-330
done
FI_B200_TRACE=1 timeout 120 python scripts/profile_step.py 512 400 f32 3 > gpurun_out/r2g_trace_step.txt 2>&1; grep "fi_b200\|^step" gpurun_out/r2g_trace_step.txt | tail -24 | cut -c 1-200
timeout 900 python -m pytest tests -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
tail -9 gpurun_out/r2g_pytest.log
timeout 400 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 900 gpurun_out/r2g_bench.json; tail -5 gpurun_out/r2g_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:stencil3d_tma_kernel --csv \
    --log-file gpurun_out/r2g_traffic.csv python scripts/profile_step.py 512 12 > gpurun_out/r2g_traffic.log 2>&1
python scripts/ncu_traffic.py gpurun_out/r2g_traffic.csv sdf3d_512_1M f32 gpurun_out/r2g_ncu_traffic.json

#!/bin/bash
# Round 2, GPU visit 5 (1 GPU): suite at HEAD (3D gradient-smoothness kernel, W-cycle defaults, vector reductions); phase trace
# of one bench step; stencil kernel variants; bench; DRAM traffic capture of the dominant kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -12 gpurun_out/r2f_pytest.log
FI_B200_TRACE=1 timeout 120 python scripts/profile_step.py 512 400 > gpurun_out/r2f_trace_step.txt 2>&1; grep "fi_b200" gpurun_out/r2f_trace_step.txt | tail -40
for v in 0 3 4; do
  FI_B200_STENCIL_VARIANT=$v timeout 200 python scripts/time_iters.py 512 1000000 200 > gpurun_out/r2f_time_iters_variant$v.jsonl 2> gpurun_out/r2f_time_iters_variant$v.err
  head -1 gpurun_out/r2f_time_iters_variant$v.jsonl | cut -c 1-330
done
timeout 300 python scripts/r2_time_to_tol.py C3 C4 512 > gpurun_out/r2f_time_to_tol.jsonl 2> gpurun_out/r2f_time_to_tol.err; cut -c 1-260 gpurun_out/r2f_time_to_tol.jsonl
timeout 400 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 1200 gpurun_out/r2f_bench.json; tail -5 gpurun_out/r2f_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:stencil3d_tma_kernel --csv \
    --log-file gpurun_out/r2f_traffic.csv python scripts/profile_step.py 512 12 > gpurun_out/r2f_traffic.log 2>&1
python scripts/ncu_traffic.py gpurun_out/r2f_traffic.csv sdf3d_512_1M f32 gpurun_out/r2f_ncu_traffic.json

#!/bin/bash
# Round 2, GPU visit 4 (1 GPU): suite with the 2D TMA kernel; W-cycle sweep; bench with all sections; DRAM traffic of the
# dominant kernel for profiles/ncu_traffic.json.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -12 gpurun_out/r2d_pytest.log
timeout 500 python scripts/r2_mg_sweep.py C3 C4 512 > gpurun_out/r2d_mg_sweep.jsonl 2> gpurun_out/r2d_mg_sweep.err; tail -3 gpurun_out/r2d_mg_sweep.err
cut -c 1-220 gpurun_out/r2d_mg_sweep.jsonl | tail -50
timeout 400 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 1500 gpurun_out/r2d_bench.json; tail -5 gpurun_out/r2d_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:stencil3d_tma_kernel --csv \
    --log-file gpurun_out/r2d_traffic.csv python scripts/profile_step.py 512 12 > gpurun_out/r2d_traffic.log 2>&1
python scripts/ncu_traffic.py gpurun_out/r2d_traffic.csv sdf3d_512_1M f32 gpurun_out/r2d_ncu_traffic.json

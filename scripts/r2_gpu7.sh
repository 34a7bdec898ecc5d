#!/bin/bash
# Round 2, GPU visit 7 (1 GPU): deferred x update, branch-free restriction, vectorised prolongation: timings, suite, bench, the
# multigrid launch list, ncu --set full captures of assembly / data term / update / 2D kernel.
set -x
mkdir -p gpurun_out
timeout 200 python scripts/time_iters.py 512,256 1000000 200 > gpurun_out/r2j_time_iters.jsonl 2> gpurun_out/r2j_time_iters.err; grep '"fast": true' gpurun_out/r2j_time_iters.jsonl | cut -c 1-330
timeout 900 python -m pytest tests -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2j_pytest.log
tail -9 gpurun_out/r2j_pytest.log
timeout 400 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 900 gpurun_out/r2j_bench.json; tail -5 gpurun_out/r2j_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2j_mg_launches.csv python scripts/profile_mg.py 512 2 f64 > gpurun_out/r2j_mg_launches.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2j_mg_launches.csv > gpurun_out/r2j_mg_launches.md 2>&1; head -16 gpurun_out/r2j_mg_launches.md | cut -c 1-200
bash scripts/r2_gpu_ncu.sh
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:stencil3d_tma_kernel --csv \
    --log-file gpurun_out/r2j_traffic.csv python scripts/profile_step.py 512 12 > gpurun_out/r2j_traffic.log 2>&1
python scripts/ncu_traffic.py gpurun_out/r2j_traffic.csv sdf3d_512_1M f32 gpurun_out/r2j_ncu_traffic.json

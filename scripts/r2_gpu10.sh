#!/bin/bash
# Round 2: the ncu launch list of the bench command itself (per-kernel share of a step) and `ncu --set full` of the multigrid
# transfer kernels and the new setup kernels.  Numbers printed under ncu are not bench values.
set -x
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 --csv \
    --log-file gpurun_out/r2r_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-time-to-tol --no-configs > gpurun_out/r2r_launches_bench.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2r_launches_bench.csv > gpurun_out/r2r_launches_bench_512_f32.md 2>&1; head -14 gpurun_out/r2r_launches_bench_512_f32.md | cut -c 1-200; tail -2 gpurun_out/r2r_launches_bench_512_f32.md
rm -f gpurun_out/r2r_launches_bench.csv
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:restrict_kernel|prolong_add4_kernel|cheb_step_kernel|mg_update_kernel|mg_direction_kernel" -c 120 \
    -o gpurun_out/r2r_mg python scripts/profile_mg.py 512 1 f64 > gpurun_out/r2r_mg_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r2r_mg.ncu-rep > gpurun_out/r2r_full_512_mg_transfers.md 2>&1; grep "^###" gpurun_out/r2r_full_512_mg_transfers.md; grep -A8 "restrict_kernel.*grid" gpurun_out/r2r_full_512_mg_transfers.md | head -12
rm -f gpurun_out/r2r_mg.ncu-rep

#!/bin/bash
# Round 2, GPU visit 2 (1 GPU): node-major data term vs the two atomic kernels; fp32 multigrid with the widening guard;
# the whole GPU suite; launch list (+ DRAM bytes) of the bench step and of two multigrid iterations.
set -x
mkdir -p gpurun_out
for v in node cell split; do
  FI_B200_DATA_TERM=$v timeout 200 python scripts/time_iters.py 512 1000000 200 > gpurun_out/r2b_time_iters_$v.jsonl 2> gpurun_out/r2b_time_iters_$v.err
done
cat gpurun_out/r2b_time_iters_*.jsonl | cut -c 1-330
timeout 300 python scripts/r2_time_to_tol.py C3 C4 512 > gpurun_out/r2b_time_to_tol.jsonl 2> gpurun_out/r2b_time_to_tol.err
cut -c 1-300 gpurun_out/r2b_time_to_tol.jsonl; tail -3 gpurun_out/r2b_time_to_tol.err
timeout 900 python -m pytest tests -m gpu -q --durations=15 -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -30 gpurun_out/r2b_pytest.log
timeout 300 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 2500 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2b_mg_launches.csv python scripts/profile_mg.py 512 2 > gpurun_out/r2b_mg_launches.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2b_mg_launches.csv > gpurun_out/r2b_mg_launches.md 2>&1; head -40 gpurun_out/r2b_mg_launches.md
rm -f gpurun_out/r2b_mg_launches.csv

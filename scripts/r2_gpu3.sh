#!/bin/bash
# Round 2, GPU visit 3 (1 GPU): suite at HEAD; multigrid tuning sweep; new bench.py (all sections); MG launch list with the
# row-blocked transfer kernels and the per-cell data term.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -12 gpurun_out/r2c_pytest.log
timeout 400 python scripts/r2_mg_sweep.py C3 C4 512 > gpurun_out/r2c_mg_sweep.jsonl 2> gpurun_out/r2c_mg_sweep.err; tail -3 gpurun_out/r2c_mg_sweep.err
cut -c 1-200 gpurun_out/r2c_mg_sweep.jsonl | tail -40
timeout 400 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 3000 gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2c_mg_launches.csv python scripts/profile_mg.py 512 2 f64 > gpurun_out/r2c_mg_launches.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2c_mg_launches.csv > gpurun_out/r2c_mg_launches.md 2>&1; head -24 gpurun_out/r2c_mg_launches.md | cut -c 1-200

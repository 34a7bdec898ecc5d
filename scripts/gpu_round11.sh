#!/bin/bash
# First GPU visit of the next round (1 GPU, ~12 min of box time): everything added after round 1's GPU budget ran out.
#   1. the whole GPU suite (the iso-surface tests have never run on a GPU; slab multigrid ran on 1 rank only)
#   2. default bench line (N=1) + reference arm
#   3. ncu launch list of the bench step and one full capture of the dominant kernels -> profiles/ summaries
# Usage: gpurun --timeout 900 -- 'bash scripts/gpu_round11.sh'
set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu11.log
tail -15 gpurun_out/pytest_gpu11.log
timeout 300 python bench.py > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -c 1500 gpurun_out/bench11.json; tail -5 gpurun_out/bench11.err
timeout 200 python bench.py --impl reference > gpurun_out/bench11_reference.json 2> gpurun_out/bench11_reference.err; tail -c 600 gpurun_out/bench11_reference.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches11.csv \
    python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu-baseline --no-time-to-tol > gpurun_out/bench11_under_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/launches11.csv > gpurun_out/launches11.md 2>&1; head -30 gpurun_out/launches11.md
# full captures (one launch of each distinct kernel): the multigrid V-cycle (TMA epilogue smoother, transfers) — the
# iteration runs at ~54 % of the HBM roofline by byte count, and nothing in profiles/ says why yet —
# and the assembly kernels (sort, scatter), which have launch-list times but no counters
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/mg11 \
    python scripts/profile_mg.py 512 1 > gpurun_out/mg11_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/mg11.ncu-rep > gpurun_out/mg11_full.md 2>&1; grep -c "^###" gpurun_out/mg11_full.md
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:scatter|radix|scan|cell_keys|canonicalise|mark_heads|slots" -c 40 -o gpurun_out/asm11 \
    python scripts/profile_step.py 512 2 > gpurun_out/asm11_ncu.log 2>&1
python scripts/ncu_summary.py full gpurun_out/asm11.ncu-rep > gpurun_out/asm11_full.md 2>&1; grep -c "^###" gpurun_out/asm11_full.md

#!/bin/bash
# Round 2, GPU visit 1 (1 GPU): the whole GPU suite at HEAD without -x, then the C3 fp32-multigrid diagnosis.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=25 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -40 gpurun_out/r2a_pytest.log
timeout 300 python scripts/r2_mg_diag.py C3 > gpurun_out/r2a_diag_c3_split.jsonl 2> gpurun_out/r2a_diag_c3_split.err
FI_B200_DATA_TERM=cell timeout 300 python scripts/r2_mg_diag.py C3 > gpurun_out/r2a_diag_c3_cell.jsonl 2> gpurun_out/r2a_diag_c3_cell.err
tail -3 gpurun_out/r2a_diag_c3_split.err
cut -c 1-400 gpurun_out/r2a_diag_c3_split.jsonl | tail -30

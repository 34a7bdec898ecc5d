#!/bin/bash
# Round 2, tail-of-cycle kernel (mg_tail_kernel): what the driver runs at round end first (whole GPU suite — it holds the
# parity test of the tail kernel —, smoke(), default bench), then what the kernel buys on the configs that fit one GPU,
# two cycle-shape variants, a phase trace and the launch list of two multigrid iterations.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2p_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2p_pytest.log
tail -5 gpurun_out/r2p_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; tail -2 gpurun_out/r2p_smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -c 600 gpurun_out/r2p_bench.json; tail -3 gpurun_out/r2p_bench.err
timeout 300 python scripts/r2_time_to_tol.py 512 C4 C3 C2 --tail=0,4096,32768 > gpurun_out/r2p_time_to_tol_tail.jsonl 2> gpurun_out/r2p_time_to_tol_tail.err
cut -c 1-330 gpurun_out/r2p_time_to_tol_tail.jsonl | grep '"rep": 2'; tail -3 gpurun_out/r2p_time_to_tol_tail.err
for kv in FI_B200_MG_WLEVELS=2 FI_B200_MG_WLEVELS=4; do
    env $kv timeout 120 python scripts/r2_time_to_tol.py 512 --tail=4096 2> /dev/null | grep '"precision": "f64"' | cut -c 1-260 | sed "s/^/$kv /" >> gpurun_out/r2p_time_to_tol_variants.txt
done
cat gpurun_out/r2p_time_to_tol_variants.txt
FI_B200_TRACE=1 timeout 120 python scripts/r2_time_to_tol.py 512 --tail=4096 > gpurun_out/r2p_trace_time_to_tol.txt 2>&1; tail -70 gpurun_out/r2p_trace_time_to_tol.txt | cut -c 1-160
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2p_mg_launches.csv python scripts/profile_mg.py 512 2 f64 > gpurun_out/r2p_mg_launches.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2p_mg_launches.csv > gpurun_out/r2p_mg_launches.md 2>&1; head -12 gpurun_out/r2p_mg_launches.md | cut -c 1-200; grep "tail\|total" gpurun_out/r2p_mg_launches.md | cut -c 1-200

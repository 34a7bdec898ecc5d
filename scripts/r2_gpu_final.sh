#!/bin/bash
# Round 2, last single-GPU visit: what the driver runs at round end — the whole GPU suite, smoke(), the default bench line —
# plus the reference arm and the DRAM-traffic capture that bench.py's roofline.traffic reads.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
tail -5 gpurun_out/r2n_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -2 gpurun_out/r2n_smoke.log
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:stencil3d_tma_kernel --csv \
    --log-file gpurun_out/r2n_traffic.csv python scripts/profile_step.py 512 12 > gpurun_out/r2n_traffic.log 2>&1
python scripts/ncu_traffic.py gpurun_out/r2n_traffic.csv sdf3d_512_1M f32 profiles/ncu_traffic.json > gpurun_out/r2n_ncu_traffic.txt; cp profiles/ncu_traffic.json gpurun_out/r2n_ncu_traffic.json
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -c 700 gpurun_out/r2n_bench.json; tail -3 gpurun_out/r2n_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err; tail -c 1200 gpurun_out/r2n_bench_reference.json

#!/bin/bash
# Round 2, last multi-GPU visit: the bench line the driver will ask for at N ranks (with trace), nothing else.
set -x
N=${1:-8}
mkdir -p gpurun_out
FI_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2o_bench_n${N}.json 2> gpurun_out/r2o_bench_n${N}.err
tail -c 3000 gpurun_out/r2o_bench_n${N}.json
for r in $(seq 0 $((N-1))); do grep "rank $r per-iteration us" gpurun_out/r2o_bench_n${N}.err | head -1; done

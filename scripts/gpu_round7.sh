#!/bin/bash
# 8-GPU visit: slab parity at 8 ranks, bench line at N=8 with the per-phase trace (run with gpurun --gpus 8).
set -x
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 scripts/slab_check.py > gpurun_out/slab_check$N.log 2>&1; echo "slab_check exit $?" >> gpurun_out/slab_check$N.log
grep -v "^\[fi\|Warning\|warn" gpurun_out/slab_check$N.log | tail -12
FI_B200_TRACE=1 timeout 200 $TR --master-port 29522 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_n$N.json; grep "per-iteration us" gpurun_out/bench_n$N.err | tail -16
grep -E "^\[fi" gpurun_out/bench_n$N.err | grep -v "per-iteration" | tail -40

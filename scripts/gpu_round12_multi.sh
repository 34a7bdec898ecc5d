#!/bin/bash
# Multi-GPU visit (gpurun --gpus 2, then --gpus 8 when the pod has room): parity of the z-slab solves against the
# single-GPU solve (Jacobi-PCG over peer memory, sharded multigrid V-cycle: validated so far only on the CPU emulator
# of tests/emu with 2-3 ranks), then the bench line with the sharded time-to-1e-6.
# Usage: gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_round12_multi.sh N'
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/slab_check.py \
    > gpurun_out/slab_check12_n$N.jsonl 2> gpurun_out/slab_check12_n$N.err; echo "exit $?" >> gpurun_out/slab_check12_n$N.jsonl
grep -v "^\[fi" gpurun_out/slab_check12_n$N.jsonl | tail -30
FI_B200_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N \
    > gpurun_out/bench12_n$N.json 2> gpurun_out/bench12_n$N.err; tail -c 2500 gpurun_out/bench12_n$N.json; grep "per-iteration us" gpurun_out/bench12_n$N.err | tail -8
# the same bench with publish / finish folded into the update kernel (opt-in; emulator-validated): two kernels fewer per iteration
FI_B200_PEER_FOLD=1 FI_B200_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --no-time-to-tol \
    > gpurun_out/bench12_fold_n$N.json 2> gpurun_out/bench12_fold_n$N.err; tail -c 1200 gpurun_out/bench12_fold_n$N.json; grep "per-iteration us" gpurun_out/bench12_fold_n$N.err | tail -8
# and on the uniform partition, to separate the two effects
FI_B200_BENCH_UNIFORM=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --no-time-to-tol \
    > gpurun_out/bench12_uniform_n$N.json 2> gpurun_out/bench12_uniform_n$N.err; tail -c 1200 gpurun_out/bench12_uniform_n$N.json
# BASELINE configs[4] (1024^3 lattice, 20M points) when all 8 GPUs are there
if [ "$N" = "8" ]; then
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --workload sdf3d_1024_20M --steps 2 --warmup 3 \
        > gpurun_out/bench12_c5_n8.json 2> gpurun_out/bench12_c5_n8.err; tail -c 2500 gpurun_out/bench12_c5_n8.json; tail -3 gpurun_out/bench12_c5_n8.err
fi

#!/bin/bash
# 2-GPU visit: slab parity vs the single-GPU solve, then the bench line at N=2 (run with gpurun --gpus 2).
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 scripts/slab_check.py > gpurun_out/slab_check2.log 2>&1; echo "slab_check exit $?" >> gpurun_out/slab_check2.log
grep -v "^\[fi\|Warning\|warn" gpurun_out/slab_check2.log | tail -40
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
FI_B200_P2P=0 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "bench(nccl path) exit $?"; tail -c 1200 gpurun_out/bench_n2_nccl.json; tail -5 gpurun_out/bench_n2_nccl.err

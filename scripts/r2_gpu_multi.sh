#!/bin/bash
# Round 2, multi-GPU visit: gpurun --gpus N -- 'bash scripts/r2_gpu_multi.sh N [full]'
#   1. the 2-rank slab parity tests (pytest)            2. who-publishes / who-finishes variants of the peer-memory PCG
#   iteration with the per-rank phase trace (FI_B200_TRACE)     3. the whole bench line at N ranks (with `full`)
set -x
N=${1:-2}
TAG=${TAG:-r2e}
FOLDS_OVERRIDE=${FOLDS_OVERRIDE:-}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_zz_multi_gpu.py tests/test_gpu_zz_slab_multigrid.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_pytest_n$N.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_n$N.log; tail -5 gpurun_out/${TAG}_pytest_n$N.log
fi
python scripts/nvlink_probe.py > gpurun_out/${TAG}_nvlink_probe_n$N.txt 2>&1
FOLDS="2 0 1 3"; if [ "$N" != "2" ]; then FOLDS="2 0"; fi; if [ -n "$FOLDS_OVERRIDE" ]; then FOLDS="$FOLDS_OVERRIDE"; fi
for fold in $FOLDS; do
  FI_B200_PEER_FOLD=$fold FI_B200_TRACE=1 run bench.py --gpus $N --steps 3 --warmup 3 --no-time-to-tol --no-configs --no-cpu-baseline \
      > gpurun_out/${TAG}_bench_n${N}_fold$fold.json 2> gpurun_out/${TAG}_bench_n${N}_fold$fold.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n${N}_fold$fold.json").read().strip().splitlines()[-1])
    print("fold $fold N $N value %.4g ms/iter %.4f" % (d["value"], d["kernels"]["slab"]["ms_per_iteration"]))
except Exception as e:
    print("fold $fold failed", e)
PY
  for r in $(seq 0 $((N-1))); do grep "rank $r per-iteration us" gpurun_out/${TAG}_bench_n${N}_fold$fold.err | tail -1; done
done
if [ "$3" = "ty8" ]; then
  FI_B200_STENCIL_TY=8 FI_B200_TRACE=1 run bench.py --gpus $N --steps 3 --warmup 3 --no-time-to-tol --no-configs --no-cpu-baseline > gpurun_out/${TAG}_bench_n${N}_ty8.json 2> gpurun_out/${TAG}_bench_n${N}_ty8.err
  tail -c 500 gpurun_out/${TAG}_bench_n${N}_ty8.json
else
  FI_B200_BENCH_UNIFORM=1 run bench.py --gpus $N --steps 3 --warmup 3 --no-time-to-tol --no-configs --no-cpu-baseline > gpurun_out/${TAG}_bench_n${N}_uniform.json 2> gpurun_out/${TAG}_bench_n${N}_uniform.err
  tail -c 600 gpurun_out/${TAG}_bench_n${N}_uniform.json
fi
if [ "$2" = "full" ]; then
  run bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n${N}_full.json 2> gpurun_out/${TAG}_bench_n${N}_full.err
  tail -c 4000 gpurun_out/${TAG}_bench_n${N}_full.json; tail -5 gpurun_out/${TAG}_bench_n${N}_full.err
fi

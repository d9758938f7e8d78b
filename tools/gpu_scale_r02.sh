#!/bin/bash
# One point of the round-2 scaling study: tools/gpu_scale_r02.sh N [weak|strong]   (run under gpurun --gpus N)
set -u
N=${1:-2}; MODE=${2:-weak}
mkdir -p gpurun_out
EXTRA=""; [ "$MODE" = strong ] && EXTRA="--strong --no-parity"
if [ "$N" = 1 ]; then
  timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_${MODE}_n1.log 2> gpurun_out/r02_scale_${MODE}_n1.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 3 --warmup 3 $EXTRA \
     > gpurun_out/r02_scale_${MODE}_n$N.log 2> gpurun_out/r02_scale_${MODE}_n$N.err
fi
echo "exit $?"; tail -c 600 gpurun_out/r02_scale_${MODE}_n$N.log

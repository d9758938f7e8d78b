#!/bin/bash
# Multi-GPU session (gpurun --gpus N): slab-mode parity test, then bench.py at 1 and N GPUs, both arms.
set -u
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_multi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpu_multi.txt 2>&1
timeout 900 python -m pytest tests/test_slab.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_slab.log 2>&1
echo "test_slab exit $?" | tee -a gpurun_out/summary_multi.txt
tail -n 5 gpurun_out/test_slab.log
if [ "${WITH_SINGLE:-1}" = "1" ]; then
  timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_n1.log 2>&1
  echo "bench n1 exit $?" | tee -a gpurun_out/summary_multi.txt
  tail -n 1 gpurun_out/bench_n1.log | cut -c 1-600
fi
for n in ${NLIST:-$N}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 3 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_n$n.log 2>&1
  echo "bench n$n exit $?" | tee -a gpurun_out/summary_multi.txt
  tail -n 1 gpurun_out/bench_n$n.log | cut -c 1-1500
done

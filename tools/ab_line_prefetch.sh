#!/bin/bash
# A/B of the next-row L2 prefetch in the warp-per-row line kernels (k_line_I, k_line_ml): per-kernel split timing of the line-Jacobi and
# multilevel iterations at 256^3 with PFEM_LINE_PREFETCH=0 / 1, then the parity tests of both preconditioners with the default (on).
set -u
mkdir -p gpurun_out
for v in 0 1; do
  PFEM_LINE_PREFETCH=$v PRE=1,2 timeout 120 python tools/time_line.py 256 012 > gpurun_out/r02_line_prefetch_$v.log 2>&1
  echo "prefetch=$v"; cat gpurun_out/r02_line_prefetch_$v.log
done
timeout 200 python -m pytest tests/test_gpu_line.py tests/test_gpu_multilevel.py -q -m gpu -p no:cacheprovider > gpurun_out/r02_line_prefetch_tests.log 2>&1
echo "tests exit $?"; tail -3 gpurun_out/r02_line_prefetch_tests.log

#!/bin/bash
set -u
mkdir -p gpurun_out
for f in tests/test_gpu_operator.py tests/test_gpu_thermal.py tests/test_gpu_shockley.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --tb=short -x -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$b exit $?"; tail -n 15 gpurun_out/$b.log | cut -c1-300
done
timeout 900 python tools/tune_fused.py 256 > gpurun_out/tune_fused.log 2>&1; echo "tune exit $?"
cat gpurun_out/tune_fused.log

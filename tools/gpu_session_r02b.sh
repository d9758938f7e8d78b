#!/bin/bash
# End-of-round-2 check on one B200 after the 2-D widening and the line-kernel prefetch: the driver's own commands (pytest -m gpu, smoke, bench).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02b_gpu.txt 2>&1
timeout -s KILL 420 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee gpurun_out/r02b_summary.txt
tail -n 3 gpurun_out/r02b_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02b_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r02b_summary.txt
tail -n 3 gpurun_out/r02b_smoke.log
timeout 300 python bench.py > gpurun_out/r02b_bench.log 2> gpurun_out/r02b_bench.err; echo "bench exit $?" | tee -a gpurun_out/r02b_summary.txt
tail -c 3000 gpurun_out/r02b_bench.log

#!/bin/bash
# ncu --set full capture of the two PCG kernels on config B (256^3) + tile sweep.
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_apply_tma|k_update' -s 6 -c 4 \
   -f -o gpurun_out/prof_pcg python bench.py --steps 1 --warmup 1 --iters 10 --no-tts --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 1200 python tools/tune_apply.py 256 0 > gpurun_out/tune.log 2>&1; echo "tune exit $?"
cat gpurun_out/tune.log

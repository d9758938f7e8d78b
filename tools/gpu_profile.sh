#!/bin/bash
# ncu --set full capture of the fused PCG kernel on config B (256^3).
set -u
mkdir -p gpurun_out
export PFEM_FUSED_TILE=${TILE:-8,2,2,3}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fpcg' -s 4 -c 2 \
   -f -o gpurun_out/prof_fused python tools/tune_fused.py 256 $PFEM_FUSED_TILE > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log

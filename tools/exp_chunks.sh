#!/bin/bash
# A/B runs of k_fpcg launch options at 256^3 (run on the GPU box)
set -u
mkdir -p gpurun_out
L=gpurun_out/exp_chunks.log
: > $L
run() { echo "== $*" >> $L; env "$@" PFEM_DEBUG_PLAN=1 timeout 300 python tools/time_line.py 256 012 2>&1 | grep -v "^$" | sort -u >> $L; }
run PRE=0 PFEM_NO_PDL=1
run PRE=0,1,2
run PRE=0 PFEM_FUSED_CHUNKS=154,51,45,6
run PRE=0 PFEM_FUSED_CHUNKS=152,60,32,12
timeout 900 python -m pytest tests/test_gpu_operator.py tests/test_gpu_thermal.py tests/test_gpu_line.py tests/test_gpu_multilevel.py tests/test_gpu_boundary.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3 >> $L
cat $L

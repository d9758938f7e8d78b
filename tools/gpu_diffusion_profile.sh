#!/bin/bash
# ncu capture of the Diffusion3D kernel (one cooperative launch = the whole loop) + smoke + adapter test; run under gpurun.
mkdir -p gpurun_out
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke2.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r02_smoke2.log
timeout -s KILL 200 python -m pytest tests/test_adapter_cpp.py -q 2>&1 | tail -3
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k k_diff_compute -c 1 -o gpurun_out/r02_prof_diff -f \
   python tools/time_diffusion.py --sizes 401 --repeat 1 --cpu-max 0 > gpurun_out/r02_ncu_diff.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r02_ncu_diff.log

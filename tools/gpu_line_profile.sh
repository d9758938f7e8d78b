#!/bin/bash
# (ncu's -k matches the function name without template arguments: the ml capture takes one whole level chain, fine level first)
# ncu --set full of the warp-per-row line kernels with the next-row L2 prefetch on (after) — the "before" is profiles/r01_line_ncu.md / r02_ml_ncu.md
set -u
mkdir -p gpurun_out
PRE=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_line_I' -s 30 -c 1 -f -o gpurun_out/r02_prof_line_pf \
   python tools/time_line.py 256 012 > gpurun_out/r02_ncu_line_pf.log 2>&1; echo "ncu line exit $?"
PRE=2 timeout 150 ncu --set full --clock-control none --import-source on -k regex:'k_line_ml' -s 30 -c 3 -f -o gpurun_out/r02_prof_ml_pf \
   python tools/time_line.py 256 012 > gpurun_out/r02_ncu_ml_pf.log 2>&1; echo "ncu ml exit $?"
ls -la gpurun_out/*.ncu-rep

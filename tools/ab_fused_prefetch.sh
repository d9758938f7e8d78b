#!/bin/bash
# A/B of the TMA L2 prefetch distance in k_fpcg (PFEM_FUSED_PREFETCH = steps beyond the shared-memory ring; 0 = off) and of the
# line-kernel prefetch distance (PFEM_LINE_PREFETCH = 1 | 2 rows ahead) at 256^3: ms per iteration of jac / ljac / mlj.
set -u
mkdir -p gpurun_out
out=gpurun_out/r02_fused_prefetch.log; : > $out
for pf in 0 1 2 4 0 2; do
  echo "PFEM_FUSED_PREFETCH=$pf" >> $out
  PFEM_FUSED_PREFETCH=$pf PRE=0,1,2 timeout 120 python tools/time_line.py 256 012 >> $out 2>&1
done
echo "PFEM_LINE_PREFETCH=2" >> $out
PFEM_LINE_PREFETCH=2 PRE=1 timeout 120 python tools/time_line.py 256 012 >> $out 2>&1
cat $out

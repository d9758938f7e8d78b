#!/bin/bash
# A/B of k_fpcg builds inside the power-capped regime of a long run: tools/ab_bench.sh lib1.so lib2.so ...  ("default" = in-tree build)
for lib in "$@"; do
  if [ "$lib" = default ]; then L=""; else L="PFEM_LIB=$PWD/$lib"; fi
  env $L python bench.py --steps 5 --warmup 3 --no-tts --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$lib', 'ms/iter %.4f' % (d['ms_per_step']/500), 'frac %.3f' % d['roofline']['frac'], 'sm_mhz', d['clocks']['sm_mhz'], 'W', d['clocks']['power_w_max'])
"
done

#!/bin/bash
# A/B timing of k_fpcg builds: tools/ab_fused.sh lib1.so lib2.so ...   (paths relative to the repo root; "default" = the in-tree build)
for lib in "$@"; do
  echo "== $lib"
  if [ "$lib" = default ]; then PRE=0 python tools/time_line.py 256 012; else PFEM_LIB=$PWD/$lib PRE=0 python tools/time_line.py 256 012; fi
done

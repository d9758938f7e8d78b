#!/bin/bash
mkdir -p gpurun_out
for c in 0 1 2 3 4 5 6 7 8 9; do ./tools/tma_probe $c; done 2>&1 | tee gpurun_out/probe.log

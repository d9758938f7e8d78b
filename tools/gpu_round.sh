#!/bin/bash
# One GPU session: parity tests (one pytest process per file so a faulting kernel cannot poison the
# rest), smoke, a short bench, and the ncu launch list.  Everything is logged under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in tests/test_gpu_multilevel.py tests/test_gpu_operator.py tests/test_gpu_thermal.py tests/test_gpu_shockley.py tests/test_gpu_thermoelectric.py tests/test_gpu_boundary.py tests/test_gpu_masked.py tests/test_gpu_line.py tests/test_gpu_layout.py tests/test_golden.py tests/test_adapter_cpp.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "$b exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$b.log
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -n 4 gpurun_out/smoke.log
if [ "${WITH_BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py --steps 3 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1; echo "bench exit $?" | tee -a gpurun_out/summary.txt
  tail -n 2 gpurun_out/bench.log
fi
if [ "${WITH_NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 2 --warmup 1 --iters 10 --no-tts --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?" | tee -a gpurun_out/summary.txt
fi
if [ "${WITH_NCU_FULL:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fpcg' -s 30 -c 2 -f -o gpurun_out/prof_fused \
     python bench.py --steps 1 --warmup 1 --iters 10 --no-tts --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?" | tee -a gpurun_out/summary.txt
  # the line-Jacobi iteration: line solve kernel + operator step (k_fpcg MODE 2)
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'k_line_I|k_fpcg<8, 2, 2, 3, 0, 2>' -s 20 -c 4 -f -o gpurun_out/prof_line \
     python tools/time_line.py 256 012 > gpurun_out/ncu_line.log 2>&1
  echo "ncu line exit $?" | tee -a gpurun_out/summary.txt
fi

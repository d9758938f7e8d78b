#!/bin/bash
# Round-2 GPU session on one B200: the driver's own commands (pytest -m gpu, smoke, bench both arms) + the ncu evidence.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout -s KILL 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee gpurun_out/r02_summary.txt
tail -n 3 gpurun_out/r02_pytest_gpu.log
timeout 900 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r02_summary.txt
tail -n 3 gpurun_out/r02_smoke.log
timeout 1200 python bench.py > gpurun_out/r02_bench.log 2> gpurun_out/r02_bench.err; echo "bench exit $?" | tee -a gpurun_out/r02_summary.txt
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_ref.log 2>&1; echo "bench ref exit $?" | tee -a gpurun_out/r02_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 2 --warmup 1 --iters 10 --no-tts --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
echo "ncu list exit $?" | tee -a gpurun_out/r02_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fpcg' -s 40 -c 1 -f -o gpurun_out/r02_prof_fpcg \
   python tools/time_line.py 256 012 > gpurun_out/r02_ncu_fpcg.log 2>&1
echo "ncu full exit $?" | tee -a gpurun_out/r02_summary.txt
# Diffusion3D: timing over mesh sizes next to the reference's band-Cholesky loop on the CPU, and one ncu capture of the cooperative kernel
timeout 600 python tools/time_diffusion.py > gpurun_out/r02_diffusion_timing.jsonl 2> gpurun_out/r02_diffusion_timing.err; echo "diffusion timing exit $?" | tee -a gpurun_out/r02_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k k_diff_compute -c 1 -f -o gpurun_out/r02_prof_diff \
   python tools/time_diffusion.py --sizes 401 --repeat 1 --cpu-max 0 > gpurun_out/r02_ncu_diff.log 2>&1
echo "ncu diffusion exit $?" | tee -a gpurun_out/r02_summary.txt

#!/bin/bash
# ncu evidence for the multilevel preconditioner: launch list of the steady-state iterations of the config B 256^3 solve
# with precond = mlj, and one --set full capture of every kernel of one iteration (level chain, k_ml_down, k_fpcg<MODE 3>).
set -u
mkdir -p gpurun_out
export PRES=mlj
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/r02_ml_launches.csv \
   python tools/time_ml.py B 256 > gpurun_out/r02_ml_ncu_list.log 2>&1
echo "ncu ml list exit $?"; tail -2 gpurun_out/r02_ml_ncu_list.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_line_ml|k_ml_down|k_fpcg' -s 50 -c 5 -f -o gpurun_out/r02_prof_ml \
   python tools/time_ml.py B 256 > gpurun_out/r02_ncu_ml.log 2>&1
echo "ncu ml full exit $?"; tail -2 gpurun_out/r02_ncu_ml.log
python tools/time_ml.py B 256 > gpurun_out/r02_ml_B_timing.jsonl 2>&1; tail -1 gpurun_out/r02_ml_B_timing.jsonl
ls -la gpurun_out/r02_prof_ml.ncu-rep

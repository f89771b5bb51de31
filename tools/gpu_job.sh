#!/bin/bash
# round-1 capture D: launch list + full captures of the production kernels after the shared first QP (k_qp0)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01d_launches.csv \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01d_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_qp0$|k_lin2|k_qp_thread' -s 1 -c 5 -f -o gpurun_out/r01d_prof \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01d_prof_run.log 2>&1
ls -la gpurun_out/

#!/bin/bash
# round-1 capture D: launch list + full captures of the production kernels (profiles/r01d_summary.md)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01d_launches.csv \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01d_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_qp0$' -s 0 -c 1 -f -o gpurun_out/r01d_prof \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01d_prof_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lin2 -s 1 -c 1 -f -o gpurun_out/r01d_lin2 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01d_lin2_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_thread -s 4 -c 1 -f -o gpurun_out/r01d_qpt \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01d_qpt_run.log 2>&1
for f in gpurun_out/r01d_prof gpurun_out/r01d_lin2 gpurun_out/r01d_qpt; do python tools/ncu_summary.py $f.ncu-rep; done

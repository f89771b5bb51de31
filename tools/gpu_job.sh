#!/bin/bash
# round-1 capture E (final code of the round): launch list + full captures of the production kernels
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01e_launches.csv \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01e_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_qp0$' -s 0 -c 1 -f -o gpurun_out/r01e_qp0 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01e_qp0_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lin2 -s 1 -c 1 -f -o gpurun_out/r01e_lin2 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01e_lin2_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_thread -s 4 -c 1 -f -o gpurun_out/r01e_qpt \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01e_qpt_run.log 2>&1
for f in gpurun_out/r01e_qp0 gpurun_out/r01e_lin2 gpurun_out/r01e_qpt; do python tools/ncu_summary.py $f.ncu-rep > $f.txt; done
python bench.py --steps 3 --warmup 3 > gpurun_out/r01e_bench_n1.json 2>/dev/null
tail -c 300 gpurun_out/r01e_bench_n1.json

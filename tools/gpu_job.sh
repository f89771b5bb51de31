#!/bin/bash
# round-1 capture D (second part): full captures of the first big k_lin2 launch (trial linearisation of all instances)
# and the first big k_qp_thread launch (second QP of every instance)
set -x
ncu --set full --clock-control none --import-source on -k regex:k_lin2 -s 1 -c 1 -f -o gpurun_out/r01d_lin2 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01d_lin2_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_thread -s 4 -c 1 -f -o gpurun_out/r01d_qpt \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01d_qpt_run.log 2>&1
ls -la gpurun_out/

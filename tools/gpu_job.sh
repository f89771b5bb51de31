#!/bin/bash
# round-1 capture C: launch list + full captures of the current production kernels, then a resident-thread sweep
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01c_launches.csv \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01c_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lin2|k_qp_thread' -s 1 -c 6 -f -o gpurun_out/r01c_prof \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r01c_prof_run.log 2>&1
for t in 256 512 1024; do
  TMPC_QP_THREADS_PER_SM=$t python bench.py --batch 262144 --steps 1 --warmup 3 --cpu-sample 1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tps $t', d['value'], d['roofline']['kernel_ms'])"
done

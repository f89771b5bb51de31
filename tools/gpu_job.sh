#!/bin/bash
# experiment: occupancy of the thread-per-instance QP kernel (register cap 128 / 80 / 64 -> 512 / 768 / 1024 threads per SM)
run() { python bench.py --steps 2 --warmup 3 --cpu-sample 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['roofline']['kernel_ms'])"; }
cp tunempc_b200/libtmpc_cstr.so /tmp/base.so
TMPC_QP_THREADS_PER_SM=512 run base512
cp tunempc_b200/exp_minb6.so tunempc_b200/libtmpc_cstr.so
TMPC_QP_THREADS_PER_SM=768 run minb6_768
cp tunempc_b200/exp_minb8.so tunempc_b200/libtmpc_cstr.so
TMPC_QP_THREADS_PER_SM=1024 run minb8_1024
TMPC_QP_THREADS_PER_SM=768 run minb8_768
cp /tmp/base.so tunempc_b200/libtmpc_cstr.so

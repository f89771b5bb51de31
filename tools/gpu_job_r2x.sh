#!/bin/bash
# round 2, capture X: resident CTAs per SM of the thread-per-instance QP kernel after the fast paths (6 / 8 / 10 / 12 x 128 threads)
mkdir -p gpurun_out
cp tunempc_b200/libtmpc_cstr.so /tmp/keep.so
for mb in 6 10 12; do
cp variants/libtmpc_cstr_minb$mb.so tunempc_b200/libtmpc_cstr.so
TMPC_QP_THREADS_PER_SM=$((mb*128)) timeout 300 python bench.py --steps 2 --warmup 3 --cpu-sample 1 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('QT_MINB=$mb  %.0f solves/s' % d['value'], {k: round(v / d['steps'], 1) for k, v in d['kernel_ms'].items()})" >> gpurun_out/r02x_minb.txt
done
cat gpurun_out/r02x_minb.txt

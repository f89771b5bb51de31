#!/bin/bash
# round 2, capture U: which variant of the terminal-row sweeps the warp kernel wants (profiling builds: a fused x2, b one per pass, c unfused)
mkdir -p gpurun_out
cp tunempc_b200/libtmpc_cstr.so /tmp/libtmpc_cstr_keep.so
for v in a b c; do
cp variants/libtmpc_cstr_prof_$v.so tunempc_b200/libtmpc_cstr.so
echo "variant $v" >> gpurun_out/r02u_phase_cycles.txt
TMPC_QP0_MIN=-1 TMPC_TRACE=1 timeout 300 python bench.py --batch 2048 --steps 1 --warmup 3 --cpu-sample 1 2>&1 >/dev/null | grep "cycles per" | tail -1 >> gpurun_out/r02u_phase_cycles.txt
timeout 300 python bench.py --batch 4096 --steps 4 --warmup 3 --cpu-sample 1 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('B=4096 %.0f solves/s' % d['value'], d['kernel_ms'])" >> gpurun_out/r02u_phase_cycles.txt
done
cat gpurun_out/r02u_phase_cycles.txt

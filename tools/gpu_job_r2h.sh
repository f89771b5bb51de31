#!/bin/bash
# round 2, capture H: thread-per-instance QP kernel variants (thread-local factorisation scratch; resident CTAs per SM)
set -x
for v in minb12:1536 minb16:2048; do
name=${v%%:*}; tps=${v#*:}
cp variants/libtmpc_cstr_$name.so tunempc_b200/libtmpc_cstr.so
TMPC_QP_THREADS_PER_SM=$tps python bench.py --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/r02h_bench_$name.json 2> gpurun_out/r02h_err_$name.log
python - <<PY
import json
d = json.load(open("gpurun_out/r02h_bench_$name.json"))
print("$name", "%.0f solves/s" % d["value"], "%.1f ms" % d["ms_per_step"], d["kernel_ms"], d["stats"]["status_hist"])
PY
done



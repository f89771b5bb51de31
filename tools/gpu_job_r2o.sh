#!/bin/bash
# round 2, capture O: warp-kernel phase cycles after packing the Schur factor back into shared memory
set -x
python -m pytest tests -m gpu -x -q -k "cstr or large_batch or edge" > gpurun_out/r02o_gputests.log 2>&1; tail -3 gpurun_out/r02o_gputests.log
for b in 4096 32768; do
python bench.py --batch $b --steps 4 --warmup 3 --cpu-sample 1 > gpurun_out/r02o_bench_b$b.json 2> gpurun_out/r02o_err.log
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02o_bench_b$b.json").read().splitlines() if l.startswith("{")][-1])
print("B=$b", "%.0f solves/s" % d["value"], "%.2f ms" % d["ms_per_step"], d["kernel_ms"], d["stats"]["status_hist"][:3])
PY
done
cp variants/libtmpc_cstr_prof.so tunempc_b200/libtmpc_cstr.so
TMPC_QP0_MIN=-1 TMPC_TRACE=1 python bench.py --batch 2048 --steps 1 --warmup 3 --cpu-sample 1 2>&1 >/dev/null | grep "cycles per" | tail -1

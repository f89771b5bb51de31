#!/bin/bash
# round 2, capture M: warp-per-instance QP kernels with the cold arrays in global memory (3 resident CTAs per SM), small batches
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02m_gputests.log 2>&1; tail -3 gpurun_out/r02m_gputests.log
for cfg in "1048576:16384" "131072:16384" "131072:200000" "32768:16384" "32768:200000" "4096:16384"; do
b=${cfg%%:*}; tm=${cfg#*:}
TMPC_QP_THREAD_MIN=$tm python bench.py --batch $b --steps 4 --warmup 3 --cpu-sample 1 > gpurun_out/r02m_bench_b${b}_tm$tm.json 2> gpurun_out/r02m_err.log
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02m_bench_b${b}_tm$tm.json").read().splitlines() if l.startswith("{")][-1])
print("B=$b thread_min=$tm", "%.0f solves/s" % d["value"], "%.2f ms" % d["ms_per_step"], d["kernel_ms"], d["stats"]["status_hist"][:3])
PY
done

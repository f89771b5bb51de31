#!/bin/bash
# round 2, capture D: threshold of the primal active-set continuation (TMPC_NONCONVEX_AFTER 2 / 3)
set -x
for na in 3 2; do
TMPC_NONCONVEX_AFTER=$na TMPC_TRACE=1 python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/r02d_bench_na$na.json 2> gpurun_out/r02d_trace_na$na.log; tail -c 700 gpurun_out/r02d_bench_na$na.json; grep "\[tmpc\]" gpurun_out/r02d_trace_na$na.log | tail -18
done
python tools/dump_stragglers.py 100 2>&1 | tail -8

#!/bin/bash
# round 2, capture A: first run of the constraint-to-go QP on the B200
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_gputests.log 2>&1; tail -5 gpurun_out/r02a_gputests.log
TMPC_TRACE=1 python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_trace.log; tail -c 1500 gpurun_out/r02a_bench.json; grep "\[tmpc\]" gpurun_out/r02a_trace.log | tail -30
TMPC_QP_MODE=w python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/r02a_bench_w.json 2>/dev/null; tail -c 600 gpurun_out/r02a_bench_w.json

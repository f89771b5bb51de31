#!/bin/bash
# round 2, capture C: primal active-set continuation only for late instances
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputests.log 2>&1; tail -5 gpurun_out/r02c_gputests.log
TMPC_TRACE=1 python bench.py --steps 2 --warmup 1 --cpu-sample 1 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_trace.log; tail -c 1500 gpurun_out/r02c_bench.json; grep "\[tmpc\]" gpurun_out/r02c_trace.log | tail -45
python tools/dump_stragglers.py 100 2>&1 | tail -5

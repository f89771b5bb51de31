#!/bin/bash
# round 2, capture F: forward/adjoint linearisation kernel k_lin3 (TMPC_LIN_MODE=3) against k_lin2
set -x
TMPC_LIN_MODE=3 python -m pytest tests -m gpu -x -q -k "cstr or unicycle or large" > gpurun_out/r02f_gputests_lin3.log 2>&1; tail -3 gpurun_out/r02f_gputests_lin3.log
for m in 3 2; do
TMPC_LIN_MODE=$m TMPC_TRACE=1 python bench.py --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/r02f_bench_lin$m.json 2> gpurun_out/r02f_trace_lin$m.log; tail -c 600 gpurun_out/r02f_bench_lin$m.json; grep "\[tmpc\] it" gpurun_out/r02f_trace_lin$m.log | tail -12 | head -6
done
TMPC_LIN_MODE=3 ncu --set full --clock-control none --import-source on -k regex:k_lin3 -s 1 -c 1 -f -o gpurun_out/r02f_lin3 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02f_lin3_run.log 2>&1
python tools/ncu_summary.py gpurun_out/r02f_lin3.ncu-rep > gpurun_out/r02f_lin3.txt; cat gpurun_out/r02f_lin3.txt

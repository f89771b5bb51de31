#!/bin/bash
# round 2, capture P: slack formulation (awe9) on the B200 -- GPU test suite, DMMA vs FMA micro-benchmark, bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_gputests.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02p_gputests.log
timeout 120 tools/dmma_bench 131072 20 > gpurun_out/r02p_dmma.json 2> gpurun_out/r02p_dmma.err
timeout 900 python bench.py --config awe9 --steps 2 --warmup 3 > gpurun_out/r02p_bench_awe9.json 2> gpurun_out/r02p_err.log
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 1 > gpurun_out/r02p_bench_cstr.json 2>> gpurun_out/r02p_err.log
tail -3 gpurun_out/r02p_gputests.log

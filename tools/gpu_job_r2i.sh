#!/bin/bash
# round 2, capture I: k_lin3 default + QP thread kernel at 8 CTAs/SM: GPU tests, bench lines, launch list, ncu --set full
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02i_gputests.log 2>&1; tail -3 gpurun_out/r02i_gputests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02i_bench_cstr.json 2> gpurun_out/r02i_bench_cstr.err; tail -c 400 gpurun_out/r02i_bench_cstr.json
for c in lq evaporation unicycle; do python bench.py --config $c --steps 2 --warmup 3 --cpu-sample 4 > gpurun_out/r02i_bench_$c.json 2> gpurun_out/r02i_bench_$c.err; tail -c 300 gpurun_out/r02i_bench_$c.json; done
python tools/bench_configs.py > gpurun_out/r02i_bench_configs.jsonl 2>&1; cat gpurun_out/r02i_bench_configs.jsonl | cut -c1-330
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02i_launches.csv \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02i_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lin3 -s 1 -c 1 -f -o gpurun_out/r02i_lin3 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02i_lin3_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_thread -s 1 -c 1 -f -o gpurun_out/r02i_qpt1 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02i_qpt1_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_qp_thread -s 3 -c 1 -f -o gpurun_out/r02i_qpt3 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02i_qpt3_run.log 2>&1
for f in gpurun_out/r02i_lin3 gpurun_out/r02i_qpt1 gpurun_out/r02i_qpt3; do python tools/ncu_summary.py $f.ncu-rep > $f.txt; done

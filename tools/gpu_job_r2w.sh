#!/bin/bash
# round 2, capture W (final code): full GPU test suite, launch list + ncu --set full of the two dominant CSTR kernels (roofline.traffic)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_gputests.log 2>&1; tail -2 gpurun_out/r02w_gputests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02w_launches.csv \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02w_launch_run.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_lin3 -s 1 -c 1 -f -o gpurun_out/r02w_lin3 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02w_lin3_run.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_qp_thread -s 3 -c 1 -f -o gpurun_out/r02w_qpt3 \
    python bench.py --batch 131072 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02w_qpt3_run.log 2>&1
for f in gpurun_out/r02w_lin3 gpurun_out/r02w_qpt3; do python tools/ncu_summary.py $f.ncu-rep > $f.txt; done
find gpurun_out -name "*.ncu-rep" -size +20M -delete
du -sh gpurun_out; ls -la gpurun_out

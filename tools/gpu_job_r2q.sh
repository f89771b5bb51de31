#!/bin/bash
# round 2, capture Q: bench lines after the unconstrained-stage fast path, ncu evidence for the DMMA / FMA comparison and for the
# awe9 kernels, awe9 with the 255-register thread kernel (variants/), awe9 at 2^16.  .ncu-rep files are summarised and dropped
# (gpurun_out/ must stay below 64 MiB).
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 1 > gpurun_out/r02q_bench_cstr.json 2> gpurun_out/r02q_err_cstr.log
for b in 131072 32768 4096; do timeout 300 python bench.py --batch $b --steps 4 --warmup 3 --cpu-sample 1 > gpurun_out/r02q_bench_b$b.json 2>> gpurun_out/r02q_err_cstr.log; done
for k in k_fma k_dmma; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 7 -c 1 -f -o gpurun_out/r02q_$k tools/dmma_bench 131072 1 > gpurun_out/r02q_${k}_run.log 2>&1
  python tools/ncu_summary.py gpurun_out/r02q_$k.ncu-rep > gpurun_out/r02q_ncu_$k.txt
done
timeout 600 python bench.py --config awe9 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02q_bench_awe9.json 2> gpurun_out/r02q_err.log
timeout 600 python bench.py --config awe9 --batch 65536 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02q_bench_awe9_b65536.json 2>> gpurun_out/r02q_err.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02q_launches_awe9.csv \
    python bench.py --config awe9 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02q_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_qp_thread -s 4 -c 1 -f -o gpurun_out/r02q_awe9_qpt \
    python bench.py --config awe9 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02q_qpt_run.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_lin3 -s 4 -c 1 -f -o gpurun_out/r02q_awe9_lin3 \
    python bench.py --config awe9 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02q_lin3_run.log 2>&1
for f in gpurun_out/r02q_awe9_qpt gpurun_out/r02q_awe9_lin3; do python tools/ncu_summary.py $f.ncu-rep > $f.txt; done
cp variants/libtmpc_awe9_minb2.so tunempc_b200/libtmpc_awe9.so
timeout 600 python bench.py --config awe9 --steps 1 --warmup 3 --cpu-sample 1 > gpurun_out/r02q_bench_awe9_minb2.json 2>> gpurun_out/r02q_err.log
find gpurun_out -name "*.ncu-rep" -size +8M -delete
du -sh gpurun_out; ls -la gpurun_out | tail -25

#!/bin/bash
# round 2, capture L (8 GPUs): weak and strong scaling lines, N = 1, 2, 4, 8
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02l_gputests.log 2>&1; tail -3 gpurun_out/r02l_gputests.log
python bench.py --steps 3 --warmup 3 --cpu-sample 1 > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_err_n1.log; tail -c 300 gpurun_out/r02l_bench_n1.json
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 --cpu-sample 1 > gpurun_out/r02l_bench_n$n.json 2> gpurun_out/r02l_err_n$n.log; tail -c 300 gpurun_out/r02l_bench_n$n.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --scaling strong --steps 3 --warmup 3 --cpu-sample 1 > gpurun_out/r02l_bench_strong_n$n.json 2> gpurun_out/r02l_err_strong_n$n.log; tail -c 300 gpurun_out/r02l_bench_strong_n$n.json
done

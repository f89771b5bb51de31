python -m pytest tests -m gpu -x -q 2>&1 | tail -5
TMPC_TRACE=1 python bench.py --batch 131072 --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/s5_tm.log 2>&1
tail -1 gpurun_out/s5_tm.log | cut -c1-200

#!/bin/bash
# round 2, capture J: small-batch throughput (strong-scaling shard sizes) with the host-driven SQP loop
set -x
for b in 131072 32768 4096; do
python bench.py --batch $b --steps 5 --warmup 3 --cpu-sample 1 > gpurun_out/r02j_bench_b$b.json 2> gpurun_out/r02j_err_b$b.log
python - <<PY
import json
d = json.load(open("gpurun_out/r02j_bench_b$b.json"))
print("B=$b", "%.0f solves/s" % d["value"], "%.2f ms" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["kernel_ms"], "launches", d["gpu_launches"])
PY
done

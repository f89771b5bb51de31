#!/bin/bash
# round 2, capture R: final code (fast path for unconstrained stages, fused terminal-row sweeps, 255-register thread kernel for wide
# stages): GPU test suite, bench lines of every config and of the small CSTR batches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02r_gputests.log 2>&1; tail -2 gpurun_out/r02r_gputests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02r_bench_cstr.json 2> gpurun_out/r02r_err.log
for b in 131072 32768 4096; do timeout 300 python bench.py --batch $b --steps 4 --warmup 3 --cpu-sample 1 > gpurun_out/r02r_bench_b$b.json 2>> gpurun_out/r02r_err.log; done
timeout 600 python bench.py --config awe9 --steps 2 --warmup 3 > gpurun_out/r02r_bench_awe9.json 2>> gpurun_out/r02r_err.log
for c in lq evaporation unicycle; do timeout 600 python bench.py --config $c --steps 2 --warmup 3 --cpu-sample 4 > gpurun_out/r02r_bench_$c.json 2>> gpurun_out/r02r_err.log; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02r_bench*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f.split('/')[-1], "%.0f solves/s" % d["value"], "e2e %.0f" % d["e2e"]["value"], "%.2f ms" % d["ms_per_step"], {k: round(v, 1) for k, v in d["kernel_ms"].items()}, d["stats"]["status_hist"][:4])
    except Exception as e:
        print(f, "ERR", e)
PY
du -sh gpurun_out

#!/bin/bash
# round 2, capture S: hybrid threshold after the warp-kernel changes, warp-kernel phase cycles (profiling build)
mkdir -p gpurun_out
for tm in 65536 200000; do for b in 131072 32768; do
TMPC_QP_THREAD_MIN=$tm timeout 300 python bench.py --batch $b --steps 4 --warmup 3 --cpu-sample 1 > gpurun_out/r02s_bench_b${b}_tm$tm.json 2>> gpurun_out/r02s_err.log; done; done
cp variants/libtmpc_cstr_prof.so tunempc_b200/libtmpc_cstr.so
for b in 2048 4096; do
TMPC_QP0_MIN=-1 TMPC_TRACE=1 timeout 300 python bench.py --batch $b --steps 1 --warmup 3 --cpu-sample 1 2>&1 >/dev/null | grep "cycles per\|qp attempts" | tail -2 > gpurun_out/r02s_phase_cycles_b$b.txt; cat gpurun_out/r02s_phase_cycles_b$b.txt; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02s_bench*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f.split('/')[-1], "%.0f solves/s" % d["value"], "%.2f ms" % d["ms_per_step"], {k: round(v, 1) for k, v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY

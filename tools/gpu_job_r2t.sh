#!/bin/bash
# round 2, capture T: pinned-input fast path of the stage elimination + register-resident fused sweeps (CSTR)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cstr or large or edge or lq" > gpurun_out/r02t_gputests.log 2>&1; tail -2 gpurun_out/r02t_gputests.log
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 1 > gpurun_out/r02t_bench_cstr.json 2> gpurun_out/r02t_err.log
for b in 131072 32768 4096; do timeout 300 python bench.py --batch $b --steps 4 --warmup 3 --cpu-sample 1 > gpurun_out/r02t_bench_b$b.json 2>> gpurun_out/r02t_err.log; done
for b in 131072 32768; do TMPC_QP_THREAD_MIN=200000 timeout 300 python bench.py --batch $b --steps 4 --warmup 3 --cpu-sample 1 > gpurun_out/r02t_bench_b${b}_warp.json 2>> gpurun_out/r02t_err.log; done
cp variants/libtmpc_cstr_prof.so tunempc_b200/libtmpc_cstr.so
TMPC_QP0_MIN=-1 TMPC_TRACE=1 timeout 300 python bench.py --batch 2048 --steps 1 --warmup 3 --cpu-sample 1 2>&1 >/dev/null | grep "cycles per\|qp attempts" | tail -2 > gpurun_out/r02t_phase_cycles_b2048.txt; cat gpurun_out/r02t_phase_cycles_b2048.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02t_bench*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f.split('/')[-1], "%.0f solves/s" % d["value"], "%.2f ms" % d["ms_per_step"], {k: round(v, 1) for k, v in d["kernel_ms"].items()}, d["stats"]["status_hist"][:3])
    except Exception as e:
        print(f, "ERR", e)
PY

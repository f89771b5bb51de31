#!/bin/bash
# round 2, capture V (8 GPUs, one process per GPU): CSTR weak and strong scaling lines, awe9 at 16 384 x0 in total and per GPU
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 "$@" > gpurun_out/r02v_$tag.json 2> gpurun_out/r02v_$tag.err; tail -c 300 gpurun_out/r02v_$tag.json | head -c 300; echo; }
run cstr_weak --steps 3 --warmup 3 --cpu-sample 1
run cstr_strong --steps 4 --warmup 3 --cpu-sample 1 --scaling strong
run awe9_strong --config awe9 --steps 2 --warmup 3 --cpu-sample 1 --scaling strong
run awe9_weak --config awe9 --steps 2 --warmup 3 --cpu-sample 1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02v_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f.split('/')[-1], d["n_gpus"], d["scaling"], "%.0f solves/s" % d["value"], "e2e %.0f" % d["e2e"]["value"], "%.2f ms" % d["ms_per_step"], d["stats"]["status_hist"][:3])
    except Exception as e:
        print(f, "ERR", e)
PY
du -sh gpurun_out

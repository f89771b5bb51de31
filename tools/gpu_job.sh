#!/bin/bash
# A/B: k_lin2 with 2 second-order pairs per consumer role (11 consumer warps, 163 registers) against 3 (7 warps, 214)
run() { python bench.py --steps 2 --warmup 3 --cpu-sample 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['stats']['status_hist'])"; }
run np3_base
cp tunempc_b200/libtmpc_cstr.so /tmp/keep.so; cp tunempc_b200/exp_np2.so tunempc_b200/libtmpc_cstr.so
run np2
cp /tmp/keep.so tunempc_b200/libtmpc_cstr.so

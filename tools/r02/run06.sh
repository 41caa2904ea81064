#!/bin/bash
# round 2, GPU call 6b: what bounds a specialised pass -- arithmetic alone (tiles cycle over an L2-resident set)
mkdir -p gpurun_out
run() { timeout 300 python bench.py --nqubit 28 --no-cpu-baseline $2 > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json; d=json.load(open('gpurun_out/tmp.json')); print('$1 [$2]', d['ms_per_step'], d['config']['passes'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['clocks'])"; tail -n 2 gpurun_out/tmp.err; }
export B200Q_JIT_PREFETCH=0
B200Q_JIT_DEBUG_ONE_TILE=1 run compute_only ""
B200Q_JIT_DEBUG_ONE_TILE=1 run compute_only "--chunk-bits 11"
B200Q_JIT_DEBUG_ONE_TILE=1 B200Q_JIT_DEBUG_SKIP_OPS=1 run l2_only ""

#!/bin/bash
# round 2, GPU call 8: two items per thread (128-thread CTAs on 64 KiB tiles, 3 CTAs per SM) vs one
mkdir -p gpurun_out
export B200Q_JIT_VERBOSE=1
run() { timeout 300 python bench.py --nqubit 28 --no-cpu-baseline $2 > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json; d=json.load(open('gpurun_out/tmp.json')); print('$1 [$2]', d['ms_per_step'], d['config']['passes'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['config']['parity_check']['rel_l2_vs_oracle_c128'], d['clocks'])"; tail -n 2 gpurun_out/tmp.err; }
run ipt1 ""
B200Q_JIT_IPT=2 run ipt2 ""
B200Q_JIT_IPT=2 B200Q_JIT_DEBUG_ONE_TILE=1 run ipt2_compute_only ""
B200Q_JIT_IPT=2 run ipt2 "--chunk-bits 13"
B200Q_JIT_IPT=2 B200Q_JIT_OVERSUB=4 run ipt2_oversub4 ""
B200Q_JIT_IPT=2 B200Q_JIT_OVERSUB=16 run ipt2_oversub16 ""

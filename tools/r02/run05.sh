#!/bin/bash
# round 2, GPU call 5: L2 prefetch of the next tile, A/B; tile size with prefetch
mkdir -p gpurun_out
export B200Q_JIT_VERBOSE=1
run() { timeout 300 python bench.py --nqubit 28 --no-cpu-baseline $2 > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json; d=json.load(open('gpurun_out/tmp.json')); print('$1 [$2]', d['ms_per_step'], d['config']['passes'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['config']['specialised_passes'], d['config']['parity_check']['rel_l2_vs_oracle_c128'], d['clocks'])"; tail -n 2 gpurun_out/tmp.err; }
run prefetch ""
run prefetch "--chunk-bits 11"
B200Q_JIT_PREFETCH=0 run noprefetch ""
B200Q_JIT_PREFETCH=0 run noprefetch "--chunk-bits 11"
timeout 400 python bench.py --nqubit 30 --no-cpu-baseline > gpurun_out/bench30_r02_d.json 2> gpurun_out/bench30_r02_d.err; cut -c1-200 gpurun_out/bench30_r02_d.json; tail -n 3 gpurun_out/bench30_r02_d.err

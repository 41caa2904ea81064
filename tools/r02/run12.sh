#!/bin/bash
# round 2, GPU call 12: per-pass efficiency of the specialised kernels on shard-sized states (1 GPU)
mkdir -p gpurun_out
run() { timeout 300 python bench.py --nqubit $2 --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json; d=json.load(open('gpurun_out/tmp.json')); print('$1 n=$2', d['ms_per_step'], d['config']['passes'], d['roofline']['frac'], d['roofline']['ms_per_launch'])"; tail -n 2 gpurun_out/tmp.err; }
run auto 27
run auto 26
run auto 25
B200Q_JIT_OVERSUB=8 run oversub8 27
B200Q_JIT_OVERSUB=1 run oversub1 27
B200Q_JIT_OVERSUB=8 run oversub8 25
B200Q_JIT_OVERSUB=1 run oversub1 25

#!/bin/bash
# round 2, GPU call 19: reference arm + default bench (the driver's commands)
mkdir -p gpurun_out
export B200Q_JIT_VERBOSE=1
SECONDS=0; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02_a.json 2> gpurun_out/bench_ref_r02_a.err; echo "ref arm: ${SECONDS}s"; cut -c1-700 gpurun_out/bench_ref_r02_a.json; tail -n 2 gpurun_out/bench_ref_r02_a.err
SECONDS=0; timeout 900 python bench.py > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err; echo "bench: ${SECONDS}s"; cut -c1-250 gpurun_out/bench_r02_d.json; tail -n 3 gpurun_out/bench_r02_d.err | cut -c1-300

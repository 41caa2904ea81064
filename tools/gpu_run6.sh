#!/bin/bash
# quick iteration: parity + bench + optional ncu of the tile kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err
if [ "$1" = "ncu" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:b200q_tile_kernel -s 60 -c 2 -o gpurun_out/prof_tile_6 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full6.log 2>&1
fi
tail -n 4 gpurun_out/pytest_gpu6.log; cut -c1-900 gpurun_out/bench6.json; tail -3 gpurun_out/bench6.err

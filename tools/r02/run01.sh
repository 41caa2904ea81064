#!/bin/bash
# round 2, GPU call 1: large-n amplitude parity of the round-1 kernel + the 30-qubit headline as a baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_large_r02_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_large_r02_a.log
tail -n 6 gpurun_out/pytest_large_r02_a.log
timeout 300 python bench.py --nqubit 30 --no-cpu-baseline > gpurun_out/bench30_r02_a.json 2> gpurun_out/bench30_r02_a.err; cut -c1-400 gpurun_out/bench30_r02_a.json; tail -n 2 gpurun_out/bench30_r02_a.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc

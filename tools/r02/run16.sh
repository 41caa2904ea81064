#!/bin/bash
# round 2, GPU call 16: first run of the tcgen05 dense block (bounded: a hang must not take the box down)
mkdir -p gpurun_out
timeout 120 python tools/dense_tc_bench.py --nqubit 26 > gpurun_out/dense_tc_r02_a.jsonl 2> gpurun_out/dense_tc_r02_a.err; echo "rc=$?"; cat gpurun_out/dense_tc_r02_a.jsonl; tail -n 5 gpurun_out/dense_tc_r02_a.err
nvidia-smi --query-gpu=name,memory.used --format=csv

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 600 python tools/microbench.py --quick > gpurun_out/microbench2.jsonl 2> gpurun_out/microbench2.err
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:b200q_tile_kernel -s 50 -c 3 -o gpurun_out/prof_tile2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -n 5 gpurun_out/pytest_gpu2.log gpurun_out/bench2.json

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err
timeout 300 python bench.py --steps 3 --warmup 3 --chunk-bits 11 --no-cpu-baseline > gpurun_out/bench4_cb11.json 2>> gpurun_out/bench4.err
timeout 600 python tools/bench_configs.py c1 c5 c3 > gpurun_out/configs4.jsonl 2> gpurun_out/configs4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:b200q_tile_kernel -s 60 -c 3 -o gpurun_out/prof_tile_r01 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full4.log 2>&1
tail -n 4 gpurun_out/pytest_gpu4.log; cut -c1-400 gpurun_out/bench4.json; cut -c1-300 gpurun_out/bench4_cb11.json; cat gpurun_out/configs4.jsonl | cut -c1-400

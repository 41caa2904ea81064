#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
timeout 600 python tools/bench_configs.py c1 c5 c3 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err
tail -n 6 gpurun_out/pytest_gpu3.log; cat gpurun_out/configs.jsonl; tail -n 5 gpurun_out/configs.err; cut -c1-300 gpurun_out/bench3.json

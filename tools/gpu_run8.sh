#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err
timeout 600 python tools/bench_configs.py c1 c5 c3 > gpurun_out/configs8.jsonl 2> gpurun_out/configs8.err
python tools/single_gate_bench.py > gpurun_out/single8.jsonl 2>&1
tail -n 4 gpurun_out/pytest_gpu8.log; cut -c1-400 gpurun_out/bench8.json; cut -c1-420 gpurun_out/configs8.jsonl; cat gpurun_out/single8.jsonl

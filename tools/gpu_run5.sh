#!/bin/bash
# GPU session 5 (re-entry): smoke, parity, bench, ncu launch list + full capture of the fused tile kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke5.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke5.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench5_ref.json 2>> gpurun_out/bench5.err
timeout 600 python tools/microbench.py > gpurun_out/microbench5.jsonl 2> gpurun_out/microbench5.err
timeout 600 python tools/bench_configs.py c1 c5 c3 > gpurun_out/configs5.jsonl 2> gpurun_out/configs5.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch5.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:b200q_tile_kernel -s 60 -c 3 -o gpurun_out/prof_tile_r01 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full5.log 2>&1
tail -n 4 gpurun_out/smoke5.log gpurun_out/pytest_gpu5.log; cut -c1-600 gpurun_out/bench5.json; cut -c1-300 gpurun_out/bench5_ref.json; cut -c1-300 gpurun_out/configs5.jsonl

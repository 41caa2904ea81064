#!/bin/bash
# GPU session 7: full parity, bench (both arms), configs C1/C3/C5, launch list + full ncu capture of the lean tile kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke7.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke7.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu7.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench7_ref.json 2>> gpurun_out/bench7.err
timeout 600 python tools/bench_configs.py c1 c5 c3 > gpurun_out/configs7.jsonl 2> gpurun_out/configs7.err
timeout 600 python tools/microbench.py > gpurun_out/microbench7.jsonl 2> gpurun_out/microbench7.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r01_d.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch7.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:b200q_tile_kernel -s 60 -c 3 -o gpurun_out/prof_tile_r01_d python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full7.log 2>&1
tail -n 3 gpurun_out/smoke7.log gpurun_out/pytest_gpu7.log; cut -c1-300 gpurun_out/bench7.json; cut -c1-400 gpurun_out/configs7.jsonl; tail -3 gpurun_out/configs7.err

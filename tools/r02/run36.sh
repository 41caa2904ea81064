#!/bin/bash
# final verification on one GPU: full -m gpu suite, smoke, default bench line, reference arm
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_final.log 2>&1; tail -2 gpurun_out/smoke_r02_final.log
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -n 1 gpurun_out/bench_r02_final.json | cut -c1-1800
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r02_final.json 2> gpurun_out/bench_ref_r02_final.err; tail -n 1 gpurun_out/bench_ref_r02_final.json | cut -c1-600

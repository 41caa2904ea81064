#!/bin/bash
# GPU session 19 (final state of the round, series r01_e): the driver's own sequence -- pytest -m gpu -x, smoke(), bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_r01_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r01_e.log
tail -n 4 gpurun_out/pytest_gpu_r01_e.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01_e.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r01_e.log; tail -n 3 gpurun_out/smoke_r01_e.log
timeout 200 python bench.py > gpurun_out/bench_r01_e.json 2> gpurun_out/bench_r01_e.err; cut -c1-330 gpurun_out/bench_r01_e.json; tail -n 2 gpurun_out/bench_r01_e.err
timeout 100 python tools/bench_denmat.py 14 > gpurun_out/denmat_final.jsonl 2>/dev/null; cut -c1-420 gpurun_out/denmat_final.jsonl

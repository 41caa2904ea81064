#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02_f.log 2>&1
tail -6 gpurun_out/pytest_gpu_r02_f.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_f.log 2>&1; tail -2 gpurun_out/smoke_r02_f.log

#!/bin/bash
# definitive record of the final commit: full -m gpu suite + smoke
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02_final3.log 2>&1
tail -3 gpurun_out/pytest_gpu_r02_final3.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

#!/bin/bash
# final verification (after the Fock folding / 29-qubit complex128 parity additions) + ncu of the staged kernel with cp.async
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02_final2.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_final2.log 2>&1; tail -2 gpurun_out/smoke_r02_final2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:qudit_sector_staged -c 1 -o gpurun_out/ncu_fock_staged_r02_b -f python tools/fock_ncu_target.py > gpurun_out/ncu_fock_staged_r02_b.log 2>&1; tail -2 gpurun_out/ncu_fock_staged_r02_b.log
timeout 300 python bench.py --config c5 > gpurun_out/bench_c5_r02_f.json 2>/dev/null; cut -c1-400 gpurun_out/bench_c5_r02_f.json

#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qudit_sector -c 2 -o gpurun_out/ncu_fock_sector_r02 -f python tools/fock_ncu_target.py > gpurun_out/ncu_fock_sector_r02.log 2>&1
tail -3 gpurun_out/ncu_fock_sector_r02.log
timeout 300 python -m pytest tests/test_widen_zfock.py -x -q -m gpu -k "native" 2>&1 | tail -3

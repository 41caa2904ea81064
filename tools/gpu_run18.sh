#!/bin/bash
# GPU session 18: full ncu capture of the qudit kernel with the staged-ELL / 4-output / f32x2 contraction
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qudit_apply -s 30 -c 1 -f -o gpurun_out/prof_qudit_r01_e2 python tools/fock_breakdown.py > gpurun_out/ncu_qudit18.log 2>&1
tail -n 2 gpurun_out/ncu_qudit18.log | cut -c1-200

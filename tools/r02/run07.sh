#!/bin/bash
# round 2, GPU call 7: ncu of the specialised kernels after the phase-gate work (28 qubits)
mkdir -p gpurun_out
export B200Q_JIT_PREFETCH=0
ncu --set full --clock-control none --import-source on -k regex:b200qj_pass -s 60 -c 4 -o gpurun_out/ncu_jit_r02_c -f python bench.py --nqubit 28 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep

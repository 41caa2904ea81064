#!/bin/bash
# round 2, GPU call 3: ncu of the specialised pass kernels (28-qubit headline circuit: same kernels, shorter replays)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02_b.csv python bench.py --nqubit 28 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
tail -n 3 gpurun_out/launches_r02_b.csv | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:b200qj_pass -s 50 -c 6 -o gpurun_out/ncu_jit_r02_b -f python bench.py --nqubit 28 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b2.log 2>&1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# GPU session 14: A/B of the Pauli-channel lowering (Bell basis vs parity blocks), launch list and one full ncu
# capture of the density-matrix workload (general kernel instantiation), reference arm of the bench
mkdir -p gpurun_out
timeout 100 python tools/bench_denmat.py 14 > gpurun_out/denmat14_bell.jsonl 2> gpurun_out/denmat14.err
B200Q_DENMAT_PAULI_BELL=0 timeout 100 python tools/bench_denmat.py 14 > gpurun_out/denmat14_parity.jsonl 2>> gpurun_out/denmat14.err
cat gpurun_out/denmat14_bell.jsonl gpurun_out/denmat14_parity.jsonl | cut -c1-900; tail -n 3 gpurun_out/denmat14.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_denmat_r01_e.csv python tools/bench_denmat.py 13 > gpurun_out/ncu_launch14.log 2>&1
grep -c tile_kernel gpurun_out/launches_denmat_r01_e.csv
timeout 60 python -m pytest tests/test_widen_denmat.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 3

#!/bin/bash
# round 2, GPU call 15: fused Fock passes after padding / 4 outputs per ELL entry
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_widen_zfock.py tests/test_fock.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_fock_r02_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fock_r02_b.log; tail -n 4 gpurun_out/pytest_fock_r02_b.log
timeout 300 python tools/bench_configs.py c5 > gpurun_out/c5_r02_b.jsonl 2> gpurun_out/c5_r02_b.err; cat gpurun_out/c5_r02_b.jsonl | cut -c1-600; tail -n 3 gpurun_out/c5_r02_b.err
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "uany or unitary or package_level or expectation" > gpurun_out/pytest_new_r02_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_r02_b.log; tail -n 4 gpurun_out/pytest_new_r02_b.log
ncu --set full --clock-control none --import-source on -k regex:qudit_fused -s 6 -c 2 -o gpurun_out/ncu_fock_fused_r02_b -f python tools/bench_configs.py c5 > gpurun_out/ncu_fock.log 2>&1; ls -la gpurun_out/ncu_fock_fused_r02_b.ncu-rep

#!/bin/bash
# round 2, GPU call 17: tensor-core block tests + ncu capture; new parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "tensor_cores or uany or unitary or package_level" > gpurun_out/pytest_tc_r02_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tc_r02_a.log; tail -n 5 gpurun_out/pytest_tc_r02_a.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_tc -s 3 -c 1 -o gpurun_out/ncu_dense_tc_r02_a -f python tools/dense_tc_bench.py --nqubit 28 > gpurun_out/ncu_tc.log 2>&1; ls -la gpurun_out/ncu_dense_tc_r02_a.ncu-rep
timeout 120 python tools/dense_tc_bench.py --nqubit 28 > gpurun_out/dense_tc_r02_b.jsonl 2>&1; tail -n 2 gpurun_out/dense_tc_r02_b.jsonl

#!/bin/bash
# GPU session 13: widened tests (density matrices, Hamiltonian, combined gate, sampled expectation) + den_mat workload
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_widen_denmat.py tests/test_widen_hamiltonian.py tests/test_widen_misc.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_widen13.log 2>&1; echo "pytest widen rc=$?" >> gpurun_out/pytest_widen13.log
tail -n 25 gpurun_out/pytest_widen13.log
timeout 120 python tools/bench_denmat.py 12 14 > gpurun_out/denmat13.jsonl 2> gpurun_out/denmat13.err; cat gpurun_out/denmat13.jsonl; tail -n 3 gpurun_out/denmat13.err

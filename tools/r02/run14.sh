#!/bin/bash
# round 2, GPU call 14: fused Fock passes, dense 5-6 target gates, one-GPU lockstep sharded test, package-level API
mkdir -p gpurun_out
export B200Q_JIT_VERBOSE=1
timeout 900 python -m pytest tests/test_widen_zfock.py tests/test_fock.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_fock_r02_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fock_r02_a.log; tail -n 6 gpurun_out/pytest_fock_r02_a.log
timeout 300 python tools/bench_configs.py c5 > gpurun_out/c5_r02_a.jsonl 2> gpurun_out/c5_r02_a.err; cat gpurun_out/c5_r02_a.jsonl | cut -c1-600; tail -n 3 gpurun_out/c5_r02_a.err
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "uany or unitary or package_level" > gpurun_out/pytest_new_r02_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new_r02_a.log; tail -n 6 gpurun_out/pytest_new_r02_a.log
timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu -p no:cacheprovider -k "one_gpu" > gpurun_out/pytest_onegpu_r02_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_onegpu_r02_a.log; tail -n 6 gpurun_out/pytest_onegpu_r02_a.log

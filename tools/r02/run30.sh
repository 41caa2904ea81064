#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_widen_zfock.py tests/test_gpu_parity.py tests/test_fock.py -x -q -m gpu -k "fock or Fock" > gpurun_out/pytest_fock_r02_i.log 2>&1
tail -5 gpurun_out/pytest_fock_r02_i.log
timeout 300 python tools/fock_gate_bench.py 2>/dev/null | grep config5
B200Q_FOCK_GROUP=0 timeout 300 python tools/fock_gate_bench.py 2>/dev/null | grep config5
timeout 300 python bench.py --config c5 > gpurun_out/bench_c5_r02_e.json 2> gpurun_out/bench_c5_r02_e.err
cat gpurun_out/bench_c5_r02_e.json; tail -3 gpurun_out/bench_c5_r02_e.err

#!/bin/bash
# structured Fock kernels: parity, per-gate timing, C5 bench line
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_widen_zfock.py tests/test_gpu_parity.py -x -q -m gpu -k "fock or Fock" > gpurun_out/pytest_fock_r02_f.log 2>&1
tail -5 gpurun_out/pytest_fock_r02_f.log
timeout 300 python tools/fock_gate_bench.py > gpurun_out/fock_gate_r02_f.jsonl 2> gpurun_out/fock_gate_r02_f.err
cat gpurun_out/fock_gate_r02_f.jsonl; tail -3 gpurun_out/fock_gate_r02_f.err
timeout 300 python tools/fock_gate_bench.py --double > gpurun_out/fock_gate_r02_f_c128.jsonl 2>> gpurun_out/fock_gate_r02_f.err
cat gpurun_out/fock_gate_r02_f_c128.jsonl
timeout 300 python bench.py --config c5 > gpurun_out/bench_c5_r02_b.json 2> gpurun_out/bench_c5_r02_b.err
cat gpurun_out/bench_c5_r02_b.json; tail -3 gpurun_out/bench_c5_r02_b.err

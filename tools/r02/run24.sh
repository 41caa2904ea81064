#!/bin/bash
# staged sector kernel for gates on the lowest mode: parity + per-gate timing + C5
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_widen_zfock.py tests/test_gpu_parity.py -x -q -m gpu -k "fock or Fock" > gpurun_out/pytest_fock_r02_h.log 2>&1
tail -5 gpurun_out/pytest_fock_r02_h.log
timeout 300 python tools/fock_gate_bench.py > gpurun_out/fock_gate_r02_h.jsonl 2> gpurun_out/fock_gate_r02_h.err
cat gpurun_out/fock_gate_r02_h.jsonl; tail -3 gpurun_out/fock_gate_r02_h.err
timeout 300 python tools/fock_gate_bench.py --double > gpurun_out/fock_gate_r02_h_c128.jsonl 2>> gpurun_out/fock_gate_r02_h.err
cat gpurun_out/fock_gate_r02_h_c128.jsonl
timeout 300 python bench.py --config c5 > gpurun_out/bench_c5_r02_d.json 2> gpurun_out/bench_c5_r02_d.err
cat gpurun_out/bench_c5_r02_d.json; tail -3 gpurun_out/bench_c5_r02_d.err

#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_widen_zfock.py tests/test_gpu_parity.py -x -q -m gpu -k "fock or Fock" 2>&1 | tail -3
timeout 200 python tools/fock_gate_bench.py --mesh 2>/dev/null | tail -1
B200Q_FOCK_FOLD=0 timeout 200 python tools/fock_gate_bench.py --mesh 2>/dev/null | tail -1
B200Q_FOCK_GROUP=0 timeout 200 python tools/fock_gate_bench.py --mesh 2>/dev/null | tail -1
timeout 200 python tools/fock_gate_bench.py 2>/dev/null | grep config5

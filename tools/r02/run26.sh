#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_widen_zfock.py tests/test_gpu_parity.py -x -q -m gpu -k "fock or Fock" 2>&1 | tail -3
for F in 32 64 128; do
  echo "F=$F"; B200Q_FOCK_STAGED_F=$F timeout 300 python tools/fock_gate_bench.py 2>/dev/null | grep -E '"modes": \[(6, 7|0, 7)\]|config5'
done
echo double; for F in 16 32 64; do
  echo "F=$F"; B200Q_FOCK_STAGED_F=$F timeout 300 python tools/fock_gate_bench.py --double 2>/dev/null | grep -E '"modes": \[(6, 7|0, 7)\]|config5'
done

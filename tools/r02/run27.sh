#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
B200Q_ADJOINT_CHUNK_BITS=11 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_widen_hamiltonian.py -x -q -m gpu -k "gradient or autograd or qaoa or dense_blocks" 2>&1 | tail -3
for cb in 0 11; do
  echo "ADJOINT_CHUNK_BITS=$cb"
  B200Q_ADJOINT_CHUNK_BITS=$cb timeout 600 python bench.py --config c3 --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['config'])"
done
timeout 300 python -m pytest tests/test_widen_qasm3.py -x -q -m gpu 2>&1 | tail -2

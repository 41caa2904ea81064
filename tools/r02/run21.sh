#!/bin/bash
# new tests of this stage: dense-block cotangents, full-size C5 photon-number checks, Hamiltonian tests
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_widen_zfock.py tests/test_widen_hamiltonian.py -x -q -m gpu \
  -k "dense_blocks or reverse_sweep_dense or full_size or hamiltonian or gradient or fock" > gpurun_out/pytest_r02_e.log 2>&1
tail -15 gpurun_out/pytest_r02_e.log

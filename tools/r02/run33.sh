#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large.py -x -q -m gpu -k "tensor_product" --durations=3 > gpurun_out/pytest_30q_parity_r02.log 2>&1
tail -12 gpurun_out/pytest_30q_parity_r02.log

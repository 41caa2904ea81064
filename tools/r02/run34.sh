#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu -k "sharded" 2>&1 | tail -3

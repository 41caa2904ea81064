#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reset.py -x -q -m gpu 2>&1 | tail -15

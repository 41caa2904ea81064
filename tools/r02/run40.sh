#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_widen_denmat.py tests/test_widen_hamiltonian.py -x -q -m gpu 2>&1 | tail -3

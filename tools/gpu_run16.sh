#!/bin/bash
# GPU session 16: qudit (Fock) kernel without per-element divisions, two outputs per thread in the contraction
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -k "fock or Fock" -p no:cacheprovider > gpurun_out/pytest_fock16.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fock16.log; tail -n 5 gpurun_out/pytest_fock16.log
timeout 100 python tools/fock_breakdown.py > gpurun_out/fock16.jsonl 2> gpurun_out/fock16.err; head -n 12 gpurun_out/fock16.jsonl; tail -n 3 gpurun_out/fock16.jsonl; tail -n 3 gpurun_out/fock16.err
timeout 100 python tools/bench_configs.py c5 > gpurun_out/configs16.jsonl 2>> gpurun_out/fock16.err; cat gpurun_out/configs16.jsonl

#!/bin/bash
# GPU session 15: A/B of the amplitude-damping lowering (rotation.diag.rotation vs general parity blocks)
mkdir -p gpurun_out
export DENMAT_ONLY_MIXED=1
timeout 100 python tools/bench_denmat.py 14 > gpurun_out/denmat15_svd.jsonl 2> gpurun_out/denmat15.err
B200Q_DENMAT_DAMPING_SVD=0 timeout 100 python tools/bench_denmat.py 14 > gpurun_out/denmat15_general.jsonl 2>> gpurun_out/denmat15.err
cat gpurun_out/denmat15_svd.jsonl gpurun_out/denmat15_general.jsonl | cut -c1-900; tail -n 3 gpurun_out/denmat15.err
timeout 60 python -m pytest tests/test_widen_denmat.py -m gpu -q -p no:cacheprovider 2>&1 | tail -n 3

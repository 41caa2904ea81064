#!/bin/bash
# GPU session 12 (short: 13 GPU-minutes left): new density-matrix / Hamiltonian tests first, then the whole GPU
# suite, the headline bench and the density-matrix workload
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_widen_denmat.py tests/test_widen_hamiltonian.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_widen12.log 2>&1; echo "pytest widen rc=$?" >> gpurun_out/pytest_widen12.log
tail -n 15 gpurun_out/pytest_widen12.log
timeout 120 python tools/bench_denmat.py 12 14 > gpurun_out/denmat12.jsonl 2> gpurun_out/denmat12.err; cat gpurun_out/denmat12.jsonl; tail -n 3 gpurun_out/denmat12.err
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/bench12.json 2> gpurun_out/bench12.err; cut -c1-700 gpurun_out/bench12.json
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu12.log
tail -n 6 gpurun_out/pytest_gpu12.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke12.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke12.log; tail -n 3 gpurun_out/smoke12.log

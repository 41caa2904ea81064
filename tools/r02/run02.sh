#!/bin/bash
# round 2, GPU call 2: first run of the per-pass specialised (NVRTC) kernels
mkdir -p gpurun_out
export B200Q_JIT_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_large_r02_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_large_r02_b.log
tail -n 6 gpurun_out/pytest_large_r02_b.log
timeout 400 python bench.py --nqubit 30 --no-cpu-baseline > gpurun_out/bench30_r02_b.json 2> gpurun_out/bench30_r02_b.err; cut -c1-900 gpurun_out/bench30_r02_b.json; tail -n 5 gpurun_out/bench30_r02_b.err
timeout 300 python bench.py --nqubit 28 --no-cpu-baseline > gpurun_out/bench28_r02_b.json 2> gpurun_out/bench28_r02_b.err; cut -c1-300 gpurun_out/bench28_r02_b.json; tail -n 5 gpurun_out/bench28_r02_b.err
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_large.py > gpurun_out/pytest_gpu_r02_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_b.log
tail -n 6 gpurun_out/pytest_gpu_r02_b.log

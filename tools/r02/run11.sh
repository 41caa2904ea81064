#!/bin/bash
# round 2, GPU call 11: 1-GPU denominator of config 4 (33 qubits depth 30), configs 1/3/5, new GPU tests
mkdir -p gpurun_out
export B200Q_JIT_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "batched_data_gradient or qaoa or autograd" > gpurun_out/pytest_grad_r02_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_grad_r02_a.log; tail -n 4 gpurun_out/pytest_grad_r02_a.log
timeout 600 python bench.py --nqubit 33 --depth 30 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_1gpu_33q_r02_a.json 2> gpurun_out/bench_1gpu_33q_r02_a.err; cut -c1-700 gpurun_out/bench_1gpu_33q_r02_a.json; tail -n 3 gpurun_out/bench_1gpu_33q_r02_a.err
timeout 900 python tools/bench_configs.py > gpurun_out/configs_r02_a.jsonl 2> gpurun_out/configs_r02_a.err; cut -c1-600 gpurun_out/configs_r02_a.jsonl; tail -n 3 gpurun_out/configs_r02_a.err

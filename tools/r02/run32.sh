#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_gpus" > gpurun_out/pytest_2gpu_r02_b.log 2>&1
tail -4 gpurun_out/pytest_2gpu_r02_b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_30q_r02_b.json 2> gpurun_out/bench_2gpu_30q_r02_b.err
tail -1 gpurun_out/bench_2gpu_30q_r02_b.json | cut -c1-700; tail -2 gpurun_out/bench_2gpu_30q_r02_b.err

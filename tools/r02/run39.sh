#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_4gpu_30q_r02_b.json 2> gpurun_out/bench_4gpu_30q_r02_b.err; tail -n 1 gpurun_out/bench_4gpu_30q_r02_b.json | cut -c1-330

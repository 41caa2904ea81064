#!/bin/bash
# 2-GPU sanity of the final code: sharded parity worker + the bench at N=2 (fused exchange)
mkdir -p gpurun_out
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29755 tests/dist_gpu_worker.py > gpurun_out/dist2_worker_r01_e.log 2>&1; echo "rc=$?" >> gpurun_out/dist2_worker_r01_e.log
tail -n 6 gpurun_out/dist2_worker_r01_e.log | cut -c1-250
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29756 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_r01_e.json 2> gpurun_out/bench_2gpu_r01_e.err
cut -c1-400 gpurun_out/bench_2gpu_r01_e.json; tail -n 2 gpurun_out/bench_2gpu_r01_e.err | cut -c1-250

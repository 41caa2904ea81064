#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/exchange_bench.py > gpurun_out/exchange2.jsonl 2> gpurun_out/exchange2.err
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_b.json 2> gpurun_out/bench_2gpu_b.err
cat gpurun_out/exchange2.jsonl; tail -n 3 gpurun_out/exchange2.err; cut -c1-1200 gpurun_out/bench_2gpu_b.json; grep -i "nvls\|channels\|P2P" gpurun_out/bench_2gpu_b.err | head -8

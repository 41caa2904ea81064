#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29755 tests/dist_gpu_worker.py > gpurun_out/dist2_worker.log 2>&1; echo "rc=$?" >> gpurun_out/dist2_worker.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29756 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_fused.json 2> gpurun_out/bench_2gpu_fused.err
B200Q_PEER_EXCHANGE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29757 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_nccl.json 2> gpurun_out/bench_2gpu_nccl.err
tail -n 12 gpurun_out/dist2_worker.log | cut -c1-300; cut -c1-1700 gpurun_out/bench_2gpu_fused.json; tail -n 4 gpurun_out/bench_2gpu_fused.err | cut -c1-300; cut -c1-400 gpurun_out/bench_2gpu_nccl.json

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29955 tests/dist_gpu_worker.py > gpurun_out/dist8_worker.log 2>&1; echo "rc=$?" >> gpurun_out/dist8_worker.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29956 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29957 bench.py --gpus 8 --steps 2 --warmup 2 --nqubit 33 --depth 30 > gpurun_out/bench_8gpu_33q.json 2> gpurun_out/bench_8gpu_33q.err
grep -E "SHARDED_OK|rc=|ok=|sharded - dense" gpurun_out/dist8_worker.log | tail -6; cut -c1-1200 gpurun_out/bench_8gpu.json; tail -n 3 gpurun_out/bench_8gpu.err | cut -c1-300; cut -c1-1200 gpurun_out/bench_8gpu_33q.json; tail -n 3 gpurun_out/bench_8gpu_33q.err | cut -c1-300

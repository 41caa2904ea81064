#!/bin/bash
# round 2, GPU call 10 (8 GPUs): sharded parity vs the oracle at 8 ranks, strong scaling at 30 qubits, config 4 (33 qubits depth 30)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tests/dist_gpu_worker.py > gpurun_out/dist8_worker_r02_a.log 2>&1; echo "worker rc=$?" >> gpurun_out/dist8_worker_r02_a.log
grep -E "n=|SHARDED|rc=" gpurun_out/dist8_worker_r02_a.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu_30q_r02_a.json 2> gpurun_out/bench_8gpu_30q_r02_a.err; tail -n 1 gpurun_out/bench_8gpu_30q_r02_a.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 8 --steps 5 --warmup 3 --nqubit 33 --depth 30 > gpurun_out/bench_8gpu_33q_r02_a.json 2> gpurun_out/bench_8gpu_33q_r02_a.err; tail -n 1 gpurun_out/bench_8gpu_33q_r02_a.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_4gpu_30q_r02_a.json 2> gpurun_out/bench_4gpu_30q_r02_a.err; tail -n 1 gpurun_out/bench_4gpu_30q_r02_a.json | cut -c1-400

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus4.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29655 tests/dist_gpu_worker.py > gpurun_out/dist4_worker.log 2>&1; echo "rc=$?" >> gpurun_out/dist4_worker.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus 4 --steps 2 --warmup 2 --nqubit 33 --depth 30 > gpurun_out/bench_4gpu_33q.json 2> gpurun_out/bench_4gpu_33q.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29658 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu_c.json 2> gpurun_out/bench_2gpu_c.err
grep -E "SHARDED_OK|rc=|ok=" gpurun_out/dist4_worker.log | tail -5; cut -c1-1300 gpurun_out/bench_4gpu.json; tail -n 3 gpurun_out/bench_4gpu.err; cut -c1-1300 gpurun_out/bench_4gpu_33q.json; tail -n 3 gpurun_out/bench_4gpu_33q.err; cut -c1-900 gpurun_out/bench_2gpu_c.json

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sharded_two_gpus" -p no:cacheprovider > gpurun_out/pytest_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 2 --warmup 2 --nqubit 31 --depth 30 > gpurun_out/bench_2gpu_31q.json 2> gpurun_out/bench_2gpu_31q.err
tail -n 8 gpurun_out/pytest_2gpu.log; cut -c1-1500 gpurun_out/bench_2gpu.json; tail -n 5 gpurun_out/bench_2gpu.err; cut -c1-1500 gpurun_out/bench_2gpu_31q.json; tail -n 5 gpurun_out/bench_2gpu_31q.err

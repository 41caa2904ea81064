#!/bin/bash
# round 2, GPU call 9 (2 GPUs): sharded path on the specialised kernels -- parity vs the oracle, bench at 30 qubits
mkdir -p gpurun_out
export B200Q_JIT_VERBOSE=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/dist_gpu_worker.py > gpurun_out/dist2_worker_r02_a.log 2>&1; echo "worker rc=$?" >> gpurun_out/dist2_worker_r02_a.log
grep -E "n=|SHARDED|rc=|Error|error" gpurun_out/dist2_worker_r02_a.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_30q_r02_a.json 2> gpurun_out/bench_2gpu_30q_r02_a.err; tail -n 1 gpurun_out/bench_2gpu_30q_r02_a.json | cut -c1-600; tail -n 3 gpurun_out/bench_2gpu_30q_r02_a.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 tools/dist_fock_gpu_check.py > gpurun_out/dist2_fock_r02_a.log 2>&1; echo "fock rc=$?" >> gpurun_out/dist2_fock_r02_a.log; tail -n 6 gpurun_out/dist2_fock_r02_a.log

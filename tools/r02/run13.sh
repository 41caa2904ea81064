#!/bin/bash
# round 2, GPU call 13 (2 GPUs): differentiable sharded expectation over NVLink + the full GPU suite on 2 GPUs
mkdir -p gpurun_out
export B200Q_JIT_VERBOSE=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/dist_gpu_worker.py > gpurun_out/dist2_worker_r02_b.log 2>&1; echo "worker rc=$?" >> gpurun_out/dist2_worker_r02_b.log
grep -E "n=|SHARDED|rc=|Error|error" gpurun_out/dist2_worker_r02_b.log | tail -14
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_r02_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_c.log; tail -n 5 gpurun_out/pytest_gpu_r02_c.log

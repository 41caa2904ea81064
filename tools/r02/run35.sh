#!/bin/bash
# 8 GPUs: sharded parity vs the oracle, then 30 q and 33 q with the segment cut (default 20) and without (0)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tests/dist_gpu_worker.py > gpurun_out/dist8_worker_r02_b.log 2>&1; echo "worker rc=$?" >> gpurun_out/dist8_worker_r02_b.log
grep -E "n=|SHARDED|rc=" gpurun_out/dist8_worker_r02_b.log | tail -6
run() {  # name, trim, extra args
  B200Q_SHARD_TRIM=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 8 --steps 5 --warmup 3 $3 > gpurun_out/$1.json 2> gpurun_out/$1.err
  tail -n 1 gpurun_out/$1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); c=d['config']; print('$1', d['ms_per_step'], d['value'], 'passes', c.get('local_passes'), 'exchanges', c.get('block_transposes'))"
}
run bench_8gpu_33q_r02_b 20 "--nqubit 33 --depth 30" 29581
run bench_8gpu_33q_r02_trim0 0 "--nqubit 33 --depth 30" 29582
run bench_8gpu_30q_r02_b 20 "" 29583
run bench_8gpu_30q_r02_trim0 0 "" 29584

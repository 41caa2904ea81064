#!/bin/bash
# round 2, GPU call 4: free phase gates (S / Z as renaming / frame bits) + A/B of tile size and grid oversubscription
mkdir -p gpurun_out
export B200Q_JIT_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_large.py -x -q -m gpu -p no:cacheprovider -k "c2_generator" > gpurun_out/pytest_large_r02_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_large_r02_c.log
tail -n 3 gpurun_out/pytest_large_r02_c.log
for cfg in "" "--chunk-bits 11" "--chunk-bits 13"; do
  timeout 300 python bench.py --nqubit 28 --no-cpu-baseline $cfg > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json; d=json.load(open('gpurun_out/tmp.json')); print('cfg [$cfg]', d['ms_per_step'], d['config']['passes'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['config']['specialised_passes'], d['config']['parity_check']['rel_l2_vs_oracle_c128'])"; tail -n 2 gpurun_out/tmp.err
done
for ov in 1 2 8; do
  B200Q_JIT_OVERSUB=$ov timeout 300 python bench.py --nqubit 28 --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err; python -c "
import json; d=json.load(open('gpurun_out/tmp.json')); print('oversub $ov', d['ms_per_step'], d['roofline']['frac'])"
done
timeout 400 python bench.py --nqubit 30 --no-cpu-baseline > gpurun_out/bench30_r02_c.json 2> gpurun_out/bench30_r02_c.err; cut -c1-200 gpurun_out/bench30_r02_c.json; tail -n 3 gpurun_out/bench30_r02_c.err

#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
for mb in 3 2; do
  echo "MIN_BLOCKS=$mb"
  B200Q_JIT_CACHE=/tmp/jc_$mb B200Q_JIT_MIN_BLOCKS=$mb timeout 900 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['config'].get('parity_check'), d['config'].get('jit'))"
done

#!/bin/bash
# round 2, GPU call 18: the driver's own sequence on the current code (pytest -m gpu, smoke, reference arm, bench) + host memory check
mkdir -p gpurun_out
free -g | head -2; nproc
export B200Q_JIT_VERBOSE=1
( /usr/bin/time -v timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02_a.json 2> gpurun_out/bench_ref_r02_a.err ); cut -c1-600 gpurun_out/bench_ref_r02_a.json; grep -E "Maximum resident|Elapsed" gpurun_out/bench_ref_r02_a.err
( /usr/bin/time -v timeout 900 python bench.py > gpurun_out/bench_r02_d.json 2> gpurun_out/bench_r02_d.err ); cut -c1-300 gpurun_out/bench_r02_d.json; grep -E "Maximum resident|Elapsed" gpurun_out/bench_r02_d.err; tail -n 3 gpurun_out/bench_r02_d.err | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_a.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02_a.log; tail -n 3 gpurun_out/smoke_r02_a.log
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_r02_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r02_d.log; tail -n 4 gpurun_out/pytest_gpu_r02_d.log

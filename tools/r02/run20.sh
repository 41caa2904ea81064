#!/bin/bash
# round 2, GPU call 20: ncu launch list + DRAM traffic of the default bench command (30 qubits); config lines c3 / c5
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_d.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
grep -c b200qj_pass gpurun_out/launches_r02_d.csv; tail -n 2 gpurun_out/launches_r02_d.csv | cut -c1-250
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:b200qj_pass -s 100 -c 6 --csv --log-file gpurun_out/traffic_r02_d.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_d2.log 2>&1
tail -n 18 gpurun_out/traffic_r02_d.csv | cut -c1-260
timeout 600 python bench.py --config c3 --steps 2 --warmup 1 > gpurun_out/bench_c3_r02_a.json 2> gpurun_out/bench_c3_r02_a.err; cut -c1-500 gpurun_out/bench_c3_r02_a.json; tail -n 2 gpurun_out/bench_c3_r02_a.err
timeout 300 python bench.py --config c5 > gpurun_out/bench_c5_r02_a.json 2> gpurun_out/bench_c5_r02_a.err; cut -c1-500 gpurun_out/bench_c5_r02_a.json; tail -n 2 gpurun_out/bench_c5_r02_a.err

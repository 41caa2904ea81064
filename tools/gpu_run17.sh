#!/bin/bash
# GPU session 17: one full ncu capture of the qudit kernel (beamsplitter, 8 modes x cutoff 10) + C5 timing
mkdir -p gpurun_out
timeout 100 python tools/bench_configs.py c5 > gpurun_out/configs17.jsonl 2> gpurun_out/configs17.err; cat gpurun_out/configs17.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qudit_apply -s 30 -c 1 -f -o gpurun_out/prof_qudit_r01_e python tools/fock_breakdown.py > gpurun_out/ncu_qudit17.log 2>&1
tail -n 3 gpurun_out/ncu_qudit17.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep

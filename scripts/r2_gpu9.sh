#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays -s 4 -c 1 -f -o /tmp/r2_ncu_trace python scripts/trace_bench.py 100 22 > gpurun_out/r2_ncu_trace.log 2>&1
ncu -i /tmp/r2_ncu_trace.ncu-rep --page raw --csv > gpurun_out/r2_ncu_trace_raw.csv 2>/dev/null
ncu -i /tmp/r2_ncu_trace.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_ncu_trace_sass.csv 2>/dev/null
ncu -i /tmp/r2_ncu_trace.ncu-rep --page details > gpurun_out/r2_ncu_trace_details.txt 2>/dev/null
tail -2 gpurun_out/r2_ncu_trace.log
timeout 300 python scripts/trace_bench.py 10 22 > gpurun_out/r2_trace_1m.json 2>&1; tail -1 gpurun_out/r2_trace_1m.json

#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --res 1024 --spp 64 --no-cpu-baseline --no-traversal > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(extend|shade|occluded|generate)' -s 48 -c 8 -o gpurun_out/prof_wave3 \
    python bench.py --steps 1 --warmup 3 --res 1024 --spp 64 --no-cpu-baseline --no-traversal > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace_rays' -s 3 -c 1 -o gpurun_out/prof_trace \
    python bench.py --steps 1 --warmup 3 --res 256 --spp 16 --no-cpu-baseline > gpurun_out/ncu_trace.log 2>&1
tail -2 gpurun_out/ncu_trace.log | cut -c1-200
ls -la gpurun_out | tail -6

#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_(extend|occluded|shade)' -s 30 -c 6 -o gpurun_out/prof_ball \
    python scripts/bench_scenes.py ball > gpurun_out/ncu_ball.log 2>&1
STRESS_INSTANCES=100 python scripts/bench_scenes.py stress 2>&1 | tail -4

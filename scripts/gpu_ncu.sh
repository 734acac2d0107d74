#!/bin/bash
# ncu captures for profiles/: launch list + full-set captures, summarised on the box (the .ncu-rep files are too big to bring back together)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --res 1024 --spp 64 --no-cpu-baseline --no-traversal --no-volpath"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(extend|shade|occluded|generate)' -s 48 -c 8 -o /tmp/prof_wave $B > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/prof_wave.ncu-rep gpurun_out/ncu_full_wavefront.json > gpurun_out/ncu_wave_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace_rays' -s 3 -c 1 -o /tmp/prof_trace \
    python bench.py --steps 1 --warmup 3 --res 256 --spp 16 --no-cpu-baseline --no-volpath > gpurun_out/ncu_trace.log 2>&1
python tools/ncu_summary.py /tmp/prof_trace.ncu-rep gpurun_out/ncu_full_trace.json > gpurun_out/ncu_trace_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_volstep' -s 6 -c 2 -o /tmp/prof_vol \
    python -c "
import sys; sys.path.insert(0, '.')
from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, smoke_scene
ctx = api.Context(0); sc = api.Scene(ctx, smoke_scene(512, 512, res=128))
sc.render(RenderParams(spp=16, rfilter='gaussian', sampler='independent', integrator='volpath'))
" > gpurun_out/ncu_vol.log 2>&1
python tools/ncu_summary.py /tmp/prof_vol.ncu-rep gpurun_out/ncu_full_volstep.json > gpurun_out/ncu_vol_summary.txt 2>&1
cp /tmp/prof_wave.ncu-rep gpurun_out/prof_wave.ncu-rep   # one raw capture travels back (26 MB)
ls -la gpurun_out | tail -12

#!/bin/bash
# ncu capture of the textured shading kernel (profiles/r01_ncu_full_textured.json)
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_shade' -s 2 -c 2 -o /tmp/prof_tex \
    python -c "
import sys; sys.path.insert(0, '.')
from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, textured_scene
ctx = api.Context(0); sc = api.Scene(ctx, textured_scene(1024, 1024, filter_type='ewa', tex_res=1024, n_theta=200, n_phi=200))
sc.render(RenderParams(spp=8, rfilter='gaussian', sampler='sobol'))
" > gpurun_out/ncu_tex.log 2>&1
timeout 100 python tools/ncu_summary.py /tmp/prof_tex.ncu-rep gpurun_out/ncu_full_textured.json > gpurun_out/ncu_tex_summary.txt 2>&1
tail -4 gpurun_out/ncu_tex.log; cat gpurun_out/ncu_tex_summary.txt | tail -12

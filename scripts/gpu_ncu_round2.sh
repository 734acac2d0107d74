#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shim.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
# ncu: one steady-state launch of each wavefront kernel (Cornell 1024^2), the traversal kernel on 10 M triangles, the volpath kernel
for k in k_generate k_extend_flat k_shade k_occluded_flat; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 1 -f -o /tmp/r2f_$k python scripts/render_once.py cornell 64 > /tmp/ncu_$k.log 2>&1
  ncu -i /tmp/r2f_$k.ncu-rep --page details > gpurun_out/r02_ncu_${k}_details.txt 2>/dev/null
done
timeout 900 ncu --set full --clock-control none -k regex:k_trace_rays -s 4 -c 1 -f -o /tmp/r2f_trace python scripts/trace_bench.py 100 22 > /tmp/ncu_trace.log 2>&1
ncu -i /tmp/r2f_trace.ncu-rep --page details > gpurun_out/r02_ncu_k_trace_rays_10M_details.txt 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:k_volstep -s 30 -c 1 -f -o /tmp/r2f_vol python scripts/render_once.py smoke 32 512 > /tmp/ncu_vol.log 2>&1
ncu -i /tmp/r2f_vol.ncu-rep --page details > gpurun_out/r02_ncu_k_volstep_lockstep_details.txt 2>/dev/null
python scripts/ncu_extract.py gpurun_out/r02_ncu_summary.json k_generate=/tmp/r2f_k_generate.ncu-rep k_extend_flat=/tmp/r2f_k_extend_flat.ncu-rep k_shade=/tmp/r2f_k_shade.ncu-rep k_occluded_flat=/tmp/r2f_k_occluded_flat.ncu-rep k_trace_rays_10M=/tmp/r2f_trace.ncu-rep k_volstep_lockstep=/tmp/r2f_vol.ncu-rep
# launch list of the bench command (shares of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02_launches_cornell.csv python scripts/render_once.py cornell 64 > /dev/null 2>&1
tail -3 gpurun_out/r02_launches_cornell.csv | cut -c1-200

#!/bin/bash
# envmap: plugin route after the Bitmap stand-in fix, A/B timings of what the map costs, ncu of the textured shading kernel under the map
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shim.py -q -p no:cacheprovider > gpurun_out/gpu_ncu_envmap_shim.log 2>&1
echo "shim pytest rc=$?" >> gpurun_out/gpu_ncu_envmap_shim.log
tail -15 gpurun_out/gpu_ncu_envmap_shim.log
for cfg in "envmap 64 1024 0" "envmap 64 1024 256" "envconst 64 1024 0" "envconst 64 1024 2" "envconst 64 1024 258" "textured 64 1024 0" "c3 64 1024 0"; do
  B2_RFILTER=gaussian python scripts/render_once.py $cfg 2>&1 | tail -1
done
B2_COMMIT_TIMING=1 B2_NINST=100 python scripts/render_once.py stress 4 512 2>&1 | grep -E "b2 commit|Msamples"
B2_RFILTER=gaussian timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_shade -s 6 -c 2 -o gpurun_out/gpu_ncu_envmap_shade python scripts/render_once.py envmap 16 512 > gpurun_out/gpu_ncu_envmap_ncu.log 2>&1
ncu -i gpurun_out/gpu_ncu_envmap_shade.ncu-rep --page details > gpurun_out/gpu_ncu_envmap_shade_details.txt 2>&1
ncu -i gpurun_out/gpu_ncu_envmap_shade.ncu-rep --page source --csv > gpurun_out/gpu_ncu_envmap_shade_source.csv 2>&1
ls -la gpurun_out/gpu_ncu_envmap_shade.ncu-rep
rm -f gpurun_out/gpu_ncu_envmap_shade.ncu-rep
head -c 4000000 gpurun_out/gpu_ncu_envmap_shade_source.csv > gpurun_out/gpu_ncu_envmap_shade_source_head.csv; rm -f gpurun_out/gpu_ncu_envmap_shade_source.csv

#!/bin/bash
# class-sorted TEX shading instances + the re-threaded host BVH build: whole GPU suite, A/B timings, commit timing, ncu of the sorted kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_env4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_env4_pytest.log
tail -25 gpurun_out/r2_env4_pytest.log
for cfg in "envmap 64 1024 0" "envmap 64 1024 256" "envconst 64 1024 0" "textured 64 1024 0" "c3 64 1024 0"; do
  B2_RFILTER=gaussian python scripts/render_once.py $cfg 2>&1 | tail -1
done
B2_COMMIT_TIMING=1 B2_NINST=100 python scripts/render_once.py stress 4 512 2>&1 | grep -E "b2 commit|Msamples"
B2_RFILTER=gaussian timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_shade -s 12 -c 3 -o gpurun_out/r2_env4_shade python scripts/render_once.py envmap 16 512 > gpurun_out/r2_env4_ncu.log 2>&1
ncu -i gpurun_out/r2_env4_shade.ncu-rep --page details > gpurun_out/r2_env4_shade_details.txt 2>&1
rm -f gpurun_out/r2_env4_shade.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
for k in k_shade k_generate k_extend k_occluded; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 1 -f -o /tmp/r2_ncu_$k python scripts/render_once.py cornell 64 > gpurun_out/r2_ncu_$k.log 2>&1
  ncu -i /tmp/r2_ncu_$k.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_ncu_$k.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_ncu_${k}_sass.csv 2>/dev/null
  ncu -i /tmp/r2_ncu_$k.ncu-rep --page details > gpurun_out/r2_ncu_${k}_details.txt 2>/dev/null
done
cp /tmp/r2_ncu_k_shade.ncu-rep gpurun_out/
du -sh gpurun_out

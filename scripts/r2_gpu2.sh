#!/bin/bash
mkdir -p gpurun_out
for v in default nofm fm_precdivsqrt ftz_only fm_triaccel; do
  if [ $v = default ]; then unset B2MTS_LIB; else export B2MTS_LIB=$PWD/mitsuba_b200/libb2mts_$v.so; fi
  timeout 300 python scripts/flip_study.py $v >> gpurun_out/r2_gpu2_flip.jsonl 2>> gpurun_out/r2_gpu2_flip.err
done
cat gpurun_out/r2_gpu2_flip.jsonl

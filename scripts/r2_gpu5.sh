#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "cornell or material_ball or pool" > gpurun_out/r2_gpu5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu5_pytest.log
tail -3 gpurun_out/r2_gpu5_pytest.log
for v in default shade5 shade4; do
  if [ $v = default ]; then unset B2MTS_LIB; else export B2MTS_LIB=$PWD/mitsuba_b200/libb2mts_$v.so; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-volpath --no-traversal --no-cpu-baseline > gpurun_out/r2_gpu5_bench_$v.json 2> gpurun_out/r2_gpu5_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_gpu5_bench_$v.json')); print('$v', round(d['value']), round(d['ms_per_step'],1), {k:round(x/3,1) for k,x in d['roofline']['kernel_ms'].items()})"
done

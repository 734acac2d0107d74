#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_gpu3_pytest.log
tail -5 gpurun_out/r2_gpu3_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-volpath --no-traversal --no-cpu-baseline > gpurun_out/r2_gpu3_bench.json 2> gpurun_out/r2_gpu3_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_gpu3_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
